"""TEST INFRASTRUCTURE ONLY.  The reference's CPU pipeline for one fragment pair, assembled from the
oracle pieces: pyramid (datasets/dataloader.py:69-189 driving the C/C++ oracle) -> KPFCNN -> losses
(trainer.py:90-98).  Used by the parity tests and timed by bench.py's cpu_baseline / --impl reference legs."""
import time

import numpy as np
import torch

from . import cpu, model_ref


def canonical_tie_order(q, s, idx):
    """Order every neighbour row by (d2, index).  The reference sorts a row by d2 with std::sort (nanoflann.hpp:1287),
    which is not stable: rows holding EXACTLY equal fp32 d2 values (3 rows of 40 000 on a 20k+20k pair) come out in an
    arbitrary order inside each equal-d2 run.  The documented tie contract of the device path (and of the plain-C port)
    is ascending index inside a run; this reorders ONLY inside runs of equal d2 (reference arithmetic, no FMA)."""
    ns = s.shape[0]
    sp = np.concatenate([s, np.full((1, 3), np.inf, np.float32)], 0)
    d = (q[:, None, :] - sp[np.minimum(idx, ns)]).astype(np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    d2 = np.where(idx >= ns, np.inf, d2).astype(np.float32)
    order = np.lexsort((idx, d2))
    return np.take_along_axis(idx, order, axis=1)


def cpu_collate(data, config, limits, impl="port", canonical_ties=True):
    """collate_fn_descriptor (dataloader.py:69-189) on the CPU oracle (`impl`: 'ref' = reference C++).
    canonical_ties: apply `canonical_tie_order` to the reference's rows BEFORE the truncation to `limits` columns
    (the port already orders ties by index)."""
    pts0, pts1, feat0, feat1, sel_corr, dist_keypts = data
    pts = np.concatenate([pts0, pts1]).astype(np.float32)
    lens = np.array([len(pts0), len(pts1)], np.int32)
    r_normal = config.first_subsampling_dl * config.conv_radius
    out = {"points": [], "neighbors": [], "pools": [], "upsamples": [], "stack_lengths": []}
    layer_blocks, layer, arch = [], 0, config.architecture

    def nb(q, s, ql, sl, r, lim):
        m = cpu.batch_query(q, s, ql, sl, r, impl=impl)
        if canonical_ties and impl == "ref" and m.shape[1] > 1:
            # chunked: the [rows, cols, 3] difference tensor of a wide (deformable-radius) matrix is large
            m = np.concatenate([canonical_tie_order(q[i:i + 8192], s, m[i:i + 8192]) for i in range(0, m.shape[0], 8192)]) \
                if m.shape[0] else m
        return torch.from_numpy(np.ascontiguousarray(m[:, :lim] if lim > 0 else m)).long()

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
    batch = cpu_collate(data, config, limits, impl=impl, canonical_ties=False)   # timed: the reference's own row order
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
