"""GPU parity of the distance / loss kernels against the reference fixtures (fwd + bwd)."""
import numpy as np
import pytest
import torch

from conftest import golden
from _util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _loss_inputs(P, rng):
    a = rng.standard_normal((P, 32)); a /= np.linalg.norm(a, axis=1, keepdims=True)
    p = a + 0.25 * rng.standard_normal((P, 32)); p /= np.linalg.norm(p, axis=1, keepdims=True)
    kp = rng.random((P, 3)) * 0.6
    dk = np.sqrt(((kp[:, None] - kp[None]) ** 2).sum(-1))
    sa, sp = rng.random((P, 1)).astype(np.float32), rng.random((P, 1)).astype(np.float32)
    return a.astype(np.float32), p.astype(np.float32), dk, sa, sp


def test_losses_vs_reference_fixture(cuda):
    from d3feat.pytorch_b200.loss import CircleLoss, ContrastiveLoss, DetLoss, PairLoss, cdist
    g = golden("losses")
    rng = np.random.default_rng(42)
    for P in (128, 64, 7):
        a, p, dk, sa, sp = _loss_inputs(P, rng)
        for kind in ("circle", "contrastive"):
            pre = "%s%d_" % (kind, P)
            for fused in (False, True):
                A = torch.from_numpy(a).to(cuda).requires_grad_(True)
                B = torch.from_numpy(p).to(cuda).requires_grad_(True)
                SA = torch.from_numpy(sa).to(cuda).requires_grad_(True)
                SP = torch.from_numpy(sp).to(cuda).requires_grad_(True)
                DK = torch.from_numpy(dk).to(cuda)
                if fused:
                    o = PairLoss(kind, "euclidean", 10, 0.1, 0.1, 1.4)(A, B, DK, SA, SP)
                    loss, det, acc, fp, an = o["desc_loss"], o["det_loss"], o["accuracy"], o["furthest_positive"], o["average_negative"]
                else:
                    mod = (CircleLoss(dist_type="euclidean", log_scale=10, safe_radius=0.1, pos_margin=0.1, neg_margin=1.4)
                           if kind == "circle" else ContrastiveLoss(0.1, 1.4, "euclidean", 0.1))
                    loss, acc, fp, an, zero, dists = mod(A, B, DK)
                    assert zero == 0 and isinstance(fp, list) and isinstance(an, list) and tuple(dists.shape) == (P, P)
                    det = DetLoss()(dists, SA, SP)
                (loss + det).backward()
                errs = dict(loss=rel_err(loss.detach().cpu(), g[pre + "loss"]), det=rel_err(det.detach().cpu(), g[pre + "det"]),
                            dA=rel_err(A.grad.cpu(), g[pre + "dA"]), dB=rel_err(B.grad.cpu(), g[pre + "dB"]),
                            dSA=rel_err(SA.grad.cpu(), g[pre + "dSA"]), dSP=rel_err(SP.grad.cpu(), g[pre + "dSP"]),
                            fp=rel_err(torch.as_tensor(fp).cpu(), g[pre + "fp"]), an=rel_err(torch.as_tensor(an).cpu(), g[pre + "an"]))
                assert abs(float(acc) - float(g[pre + "acc"])) < 1e-3
                assert max(errs.values()) < TOL, (kind, P, fused, errs)
        if P == 64:
            for metric in ("cosine", "sqeuclidean", "cityblock", "arccosine"):
                d = cdist(torch.from_numpy(a).to(cuda), torch.from_numpy(p).to(cuda), metric).cpu().numpy()
                assert rel_err(d, g["cdist_" + metric]) < TOL, metric


@pytest.mark.parametrize("metric", ["euclidean", "sqeuclidean", "cityblock", "cosine"])
@pytest.mark.parametrize("kind", ["circle", "contrastive"])
def test_loss_gradients_other_metrics_vs_oracle(cuda, metric, kind):
    from oracle import model_ref
    from d3feat.pytorch_b200.loss import PairLoss
    rng = np.random.default_rng(7)
    a, p, dk, sa, sp = _loss_inputs(48, rng)
    A = torch.from_numpy(a).requires_grad_(True); B = torch.from_numpy(p).requires_grad_(True)
    SA = torch.from_numpy(sa).requires_grad_(True); SP = torch.from_numpy(sp).requires_grad_(True)
    if kind == "circle":
        l, _, _, _, d = model_ref.circle_loss(A, B, torch.from_numpy(dk), dist_type=metric)
    else:
        l, _, _, _, d = model_ref.contrastive_loss(A, B, torch.from_numpy(dk), metric=metric, safe_radius=0.1)
    det = model_ref.det_loss(d, SA, SP)
    (l + 0.5 * det).backward()
    Ag = torch.from_numpy(a).to(cuda).requires_grad_(True); Bg = torch.from_numpy(p).to(cuda).requires_grad_(True)
    SAg = torch.from_numpy(sa).to(cuda).requires_grad_(True); SPg = torch.from_numpy(sp).to(cuda).requires_grad_(True)
    o = PairLoss(kind, metric, 10, 0.1, 0.1, 1.4)(Ag, Bg, torch.from_numpy(dk).to(cuda), SAg, SPg)
    (o["desc_loss"] + 0.5 * o["det_loss"]).backward()
    errs = dict(l=rel_err(o["desc_loss"].detach().cpu(), l.detach()), det=rel_err(o["det_loss"].detach().cpu(), det.detach()),
                dA=rel_err(Ag.grad.cpu(), A.grad), dB=rel_err(Bg.grad.cpu(), B.grad), dS=rel_err(SAg.grad.cpu(), SA.grad))
    assert max(errs.values()) < TOL, errs


@pytest.mark.parametrize("P", [1, 7, 300])
def test_det_loss_on_an_arbitrary_distance_matrix_vs_oracle(cuda, P):
    """DetLoss must accept ANY [P,P] matrix, as the reference does (utils/loss.py:149-158): a clone / cdist output has
    lost the tag of this package's CircleLoss and takes the stand-alone kernels; value and all three gradients are compared
    with the oracle's detector loss.  One diagonal entry is negative (furthest positive = 0, no gradient to it)."""
    from oracle import model_ref
    from d3feat.pytorch_b200.loss import DetLoss
    rng = np.random.default_rng(P)
    d = (rng.random((P, P)) * 2).astype(np.float32)
    if P > 1:
        d[P // 2, P // 2] = -0.3
    sa, sp = rng.random((P, 1)).astype(np.float32), rng.random((P, 1)).astype(np.float32)
    D = torch.from_numpy(d).requires_grad_(True); SA = torch.from_numpy(sa).requires_grad_(True); SP = torch.from_numpy(sp).requires_grad_(True)
    ref = model_ref.det_loss(D, SA, SP)
    (3.0 * ref).backward()
    Dg = torch.from_numpy(d).to(cuda).requires_grad_(True)
    SAg = torch.from_numpy(sa).to(cuda).requires_grad_(True); SPg = torch.from_numpy(sp).to(cuda).requires_grad_(True)
    out = DetLoss()(Dg.clone(), SAg, SPg)
    (3.0 * out).backward()
    assert rel_err(out.detach().cpu(), ref.detach()) < TOL
    assert rel_err(SAg.grad.cpu(), SA.grad) < TOL and rel_err(SPg.grad.cpu(), SP.grad) < TOL
    assert rel_err(Dg.grad.cpu(), D.grad) < TOL
    with pytest.raises(RuntimeError):
        DetLoss()(torch.from_numpy(d), SA, SP)          # CPU tensors: no CPU path


@pytest.mark.parametrize("W,P,D,dk_dtype", [(3, 7, 32, torch.float64), (2, 128, 32, torch.float64), (8, 16, 5, torch.float32)])
def test_exchange_pack_unpack_bit_exact(cuda, W, P, D, dk_dtype):
    """d3f_exchange_pack / d3f_exchange_unpack (the multi-GPU exchange step around the all-gather): W packed chunks
    unpack to exactly the rows that went in, with dist_keypts block-diagonal and +inf between different fragments."""
    from d3feat.pytorch_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(W * 100 + P)
    st = torch.cuda.current_stream().cuda_stream
    chunk = int(lib.d3f_exchange_chunk_bytes(P, D))
    assert chunk == (2 * P * D + 2 * P) * 4 + P * P * 8
    allbuf = torch.empty(W * chunk, dtype=torch.uint8, device=cuda)
    parts = []
    for r in range(W):
        a, p = torch.randn(P, D, generator=g).to(cuda), torch.randn(P, D, generator=g).to(cuda)
        sa, sp = torch.rand(P, generator=g).to(cuda), torch.rand(P, generator=g).to(cuda)
        dk = torch.rand(P, P, generator=g, dtype=torch.float64).to(dk_dtype).to(cuda)
        parts.append((a, p, sa, sp, dk))
        _lib.check(lib.d3f_exchange_pack(a.data_ptr(), p.data_ptr(), sa.data_ptr(), sp.data_ptr(), dk.data_ptr(),
                                         int(dk_dtype == torch.float64), P, D, allbuf[r * chunk:].data_ptr(), st))
    A = torch.empty(W * P, D, device=cuda); Pos = torch.empty_like(A)
    SA = torch.empty(W * P, device=cuda); SP = torch.empty_like(SA)
    DK = torch.empty(W * P, W * P, dtype=torch.float64, device=cuda)
    _lib.check(lib.d3f_exchange_unpack(allbuf.data_ptr(), W, P, D, A.data_ptr(), Pos.data_ptr(), SA.data_ptr(), SP.data_ptr(),
                                       DK.data_ptr(), st))
    want = torch.full((W * P, W * P), float("inf"), dtype=torch.float64, device=cuda)
    for r, (a, p, sa, sp, dk) in enumerate(parts):
        want[r * P:(r + 1) * P, r * P:(r + 1) * P] = dk.double()
    assert torch.equal(A, torch.cat([t[0] for t in parts])) and torch.equal(Pos, torch.cat([t[1] for t in parts]))
    assert torch.equal(SA, torch.cat([t[2] for t in parts])) and torch.equal(SP, torch.cat([t[3] for t in parts]))
    assert torch.equal(DK, want)
