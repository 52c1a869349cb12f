"""TEST INFRASTRUCTURE ONLY.  The reference's CPU pipeline for one fragment pair, assembled from the
oracle pieces: pyramid (datasets/dataloader.py:69-189 driving the C/C++ oracle) -> KPFCNN -> losses
(trainer.py:90-98).  Used by the parity tests and timed by bench.py's cpu_baseline / --impl reference legs."""
import time

import numpy as np
import torch

from . import cpu, model_ref


def cpu_collate(data, config, limits, impl="port"):
    """collate_fn_descriptor (dataloader.py:69-189) on the CPU oracle (`impl`: 'ref' = reference C++)."""
    pts0, pts1, feat0, feat1, sel_corr, dist_keypts = data
    pts = np.concatenate([pts0, pts1]).astype(np.float32)
    lens = np.array([len(pts0), len(pts1)], np.int32)
    r_normal = config.first_subsampling_dl * config.conv_radius
    out = {"points": [], "neighbors": [], "pools": [], "upsamples": [], "stack_lengths": []}
    layer_blocks, layer, arch = [], 0, config.architecture

    def nb(q, s, ql, sl, r, lim):
        m = cpu.batch_query(q, s, ql, sl, r, impl=impl)
        return torch.from_numpy(m[:, :lim] if lim > 0 else m).long()

    for bi, block in enumerate(arch):
        if "global" in block or "upsample" in block:
            break
        if not ("pool" in block or "strided" in block):
            layer_blocks.append(block)
            if bi < len(arch) - 1 and "upsample" not in arch[bi + 1]:
                continue
        if layer_blocks:
            deform = any("deformable" in b for b in layer_blocks[:-1])
            r = r_normal * config.deform_radius / config.conv_radius if deform else r_normal
            conv_i = nb(pts, pts, lens, lens, r, limits[layer])
        else:
            conv_i = torch.zeros((0, 1), dtype=torch.int64)
        if "pool" in block or "strided" in block:
            dl = 2 * r_normal / config.conv_radius
            pool_p, pool_b = cpu.subsample_batch(pts, lens, dl, impl=impl)
            r = r_normal * config.deform_radius / config.conv_radius if "deformable" in block else r_normal
            pool_i = nb(pool_p, pts, pool_b, lens, r, limits[layer])
            up_i = nb(pts, pool_p, lens, pool_b, 2 * r, limits[layer])
        else:
            pool_i = up_i = torch.zeros((0, 1), dtype=torch.int64)
            pool_p, pool_b = np.zeros((0, 3), np.float32), np.zeros((0,), np.int32)
        out["points"].append(torch.from_numpy(pts))
        out["neighbors"].append(conv_i)
        out["pools"].append(pool_i)
        out["upsamples"].append(up_i)
        out["stack_lengths"].append(torch.from_numpy(lens))
        pts, lens = pool_p, pool_b
        r_normal *= 2
        layer += 1
        layer_blocks = []
    out["features"] = torch.from_numpy(np.concatenate([feat0, feat1]).astype(np.float32))
    out["corr"] = torch.from_numpy(sel_corr)
    out["dist_keypts"] = torch.from_numpy(dist_keypts)
    return out


def cpu_pair_step(data, sd, config, limits, impl="ref", backward=True, sgd=None):
    """One whole pair on the CPU: collate -> forward -> circle + detector loss (-> backward -> SGD step,
    training_3DMatch.py:62-69: lr 0.01, momentum 0.98, weight decay 1e-6; `sgd` = dict of momentum buffers,
    updated in place together with `sd`).  Returns (seconds per stage dict, loss value)."""
    t = {}
    t0 = time.perf_counter()
    batch = cpu_collate(data, config, limits, impl=impl)
    t["collate"] = time.perf_counter() - t0
    params = {k: (v.detach().clone().requires_grad_(backward and "kernel_points" not in k)) for k, v in sd.items()}
    t0 = time.perf_counter()
    with torch.set_grad_enabled(backward):
        f, s = model_ref.kpfcnn_forward(params, batch, config, training=True)
        t["forward"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        dl, det, acc, _ = model_ref.pair_losses(f, s, batch, "circle")
        loss = dl + det
        t["loss"] = time.perf_counter() - t0
        if backward:
            t0 = time.perf_counter()
            loss.backward()
            t["backward"] = time.perf_counter() - t0
            if sgd is not None:
                t0 = time.perf_counter()
                with torch.no_grad():
                    for k, prm in params.items():
                        if prm.grad is None:
                            continue
                        g = prm.grad + 1e-6 * prm
                        buf = sgd.get(k)
                        buf = g.clone() if buf is None else buf.mul_(0.98).add_(g)
                        sgd[k] = buf
                        sd[k] = (prm - 0.01 * buf).detach()
                t["sgd"] = time.perf_counter() - t0
    return t, float(loss.detach())
