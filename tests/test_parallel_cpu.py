"""world_size-2 gloo tests (CPU) of the multi-GPU host logic (d3feat/pytorch_b200/parallel.py): packed
all-gather, block-diagonal keypoint distances, local-slice gradients, flat gradient all-reduce.
The GPU loss kernel is replaced by the oracle's circle/detector loss (tests may use the oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs(rank, P=16, D=8):
    rng = np.random.default_rng(100 + rank)
    a = rng.standard_normal((P, D)); a /= np.linalg.norm(a, axis=1, keepdims=True)
    p = a + 0.3 * rng.standard_normal((P, D)); p /= np.linalg.norm(p, axis=1, keepdims=True)
    kp = rng.random((P, 3)) * 0.5
    dk = np.sqrt(((kp[:, None] - kp[None]) ** 2).sum(-1))
    return (torch.from_numpy(a.astype(np.float32)), torch.from_numpy(p.astype(np.float32)),
            torch.from_numpy(rng.random((P, 1)).astype(np.float32)), torch.from_numpy(rng.random((P, 1)).astype(np.float32)),
            torch.from_numpy(dk))


def _oracle_loss(A, Pos, DK, SA, SP):
    from oracle import model_ref
    dl, acc, fp, an, d = model_ref.circle_loss(A, Pos, DK)
    return {"desc_loss": dl, "det_loss": model_ref.det_loss(d, SA, SP)}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from d3feat.pytorch_b200 import parallel
    a, p, sa, sp, dk = _inputs(rank)
    a.requires_grad_(True); p.requires_grad_(True); sa.requires_grad_(True); sp.requires_grad_(True)
    o = parallel.cross_fragment_loss(_oracle_loss, a, p, dk, sa, sp)
    (o["desc_loss"] + o["det_loss"]).backward()
    # flat gradient all-reduce on a toy module
    lin = torch.nn.Linear(4, 3)
    with torch.no_grad():
        for q in lin.parameters():
            q.fill_(0.5)
    fg = parallel.FlatGradients(lin)
    fg.zero()
    lin(torch.full((2, 4), float(rank + 1))).sum().backward()
    fg.allreduce()
    out[rank] = dict(desc=float(o["desc_loss"]), det=float(o["det_loss"]), ga=a.grad.clone(), gp=p.grad.clone(),
                     gsa=sa.grad.clone(), wgrad=lin.weight.grad.clone(), flat=fg.flat.clone())
    dist.destroy_process_group()


def test_cross_fragment_loss_world2_matches_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    # single-process reference: concatenated batch, block-diagonal dist_keypts with +inf off the blocks
    parts = [_inputs(r) for r in range(world)]
    A = torch.cat([x[0] for x in parts]).requires_grad_(True)
    Pos = torch.cat([x[1] for x in parts]).requires_grad_(True)
    SA = torch.cat([x[2] for x in parts]).requires_grad_(True)
    SP = torch.cat([x[3] for x in parts]).requires_grad_(True)
    P = parts[0][0].shape[0]
    DK = torch.full((world * P, world * P), float("inf"), dtype=torch.float64)
    for r in range(world):
        DK[r * P:(r + 1) * P, r * P:(r + 1) * P] = parts[r][4]
    o = _oracle_loss(A, Pos, DK, SA, SP)
    (o["desc_loss"] + o["det_loss"]).backward()
    for r in range(world):
        assert abs(out[r]["desc"] - float(o["desc_loss"])) < 1e-6 and abs(out[r]["det"] - float(o["det_loss"])) < 1e-6
        assert torch.allclose(out[r]["ga"], A.grad[r * P:(r + 1) * P], atol=1e-7)     # each rank owns its slice of the gradient
        assert torch.allclose(out[r]["gp"], Pos.grad[r * P:(r + 1) * P], atol=1e-7)
        assert torch.allclose(out[r]["gsa"], SA.grad[r * P:(r + 1) * P], atol=1e-7)
        # SUM all-reduce of d/dW sum(lin(x)) with x = rank+1: (1 + 2) * 2 rows
        assert torch.allclose(out[r]["wgrad"], torch.full((3, 4), 6.0))
        assert out[r]["flat"].numel() == 15


def test_single_rank_is_identity():
    from d3feat.pytorch_b200 import parallel
    a, p, sa, sp, dk = _inputs(0)
    o1 = parallel.cross_fragment_loss(_oracle_loss, a, p, dk, sa, sp)
    o2 = _oracle_loss(a, p, dk, sa, sp)
    assert float(o1["desc_loss"]) == float(o2["desc_loss"]) and float(o1["det_loss"]) == float(o2["det_loss"])


def test_pack_roundtrip():
    from d3feat.pytorch_b200.parallel import _pack, _unpack
    a, p, sa, sp, dk = _inputs(3, P=5, D=32)
    ua, up, usa, usp, udk = _unpack(_pack(a, p, sa, sp, dk), 5, 32)
    assert torch.equal(ua, a) and torch.equal(up, p) and torch.equal(usa, sa) and torch.equal(usp, sp) and torch.equal(udk, dk)


def test_attach_local_rows_gradient_is_the_local_slice():
    """parallel._AttachLocal (the CUDA exchange path's replacement for torch.cat([const.., local, const..])): forward is
    the gathered matrix itself, backward hands the local rows' gradient to the local tensor and nothing to the rest."""
    from d3feat.pytorch_b200 import parallel
    g = torch.Generator().manual_seed(3)
    local = torch.randn(4, 5, generator=g, requires_grad=True)
    others = [torch.randn(4, 5, generator=g) for _ in range(2)]
    w = torch.randn(12, 5, generator=g)
    want = torch.cat([others[0], local, others[1]])
    (want * w).sum().backward()
    ref_grad, local.grad = local.grad.clone(), None
    gathered = torch.cat([others[0], local.detach(), others[1]])
    out = parallel._AttachLocal.apply(local, gathered, 4)
    assert out.requires_grad and torch.equal(out.detach(), want.detach())
    (out * w).sum().backward()
    assert torch.equal(local.grad, ref_grad)
