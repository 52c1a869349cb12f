"""TEST INFRASTRUCTURE ONLY.  Torch-CPU restatement of the floating-point half of the
D3Feat hot path: KPConv (rigid / deformable / modulated), the KPFCNN harness around
it, and the descriptor / detector losses.  Functional style, driven by a reference
``state_dict`` so parity tests can feed the *same* parameters to this oracle and to
the CUDA product path.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified reference
(models/blocks.py, models/architectures.py, utils/loss.py) in the build container,
checks every function below against it on seeded inputs (fwd and autograd bwd), and
writes the fixtures in tests/golden/ that tests/test_oracle_cpu.py re-checks without
the reference tree.

Everything is computed with stock torch ops in the dtype of the inputs (fp32 for
parity runs, fp64 to arbitrate 1e-4 disputes).  Reference file:line cited per function.
"""
import math

import torch
import torch.nn.functional as F

SHADOW_COORD = 1e6  # models/blocks.py:277


# --------------------------------------------------------------------------- KPConv
def kp_influence(sq_d, extent, mode):
    """blocks.py:329-343.  sq_d [..] squared distance to a kernel point -> weight."""
    if mode == "linear":
        return torch.clamp(1.0 - torch.sqrt(sq_d) / extent, min=0.0)
    if mode == "constant":
        return torch.ones_like(sq_d)
    if mode == "gaussian":
        sigma = extent * 0.3
        return torch.exp(-sq_d / (2 * sigma ** 2 + 1e-9))  # blocks.py:69-76
    raise ValueError(mode)


def kpconv_rigid(q_pts, s_pts, inds, x, weights, kernel_points, extent,
                 influence="linear", aggregation="sum", kp_per_query=None, modulations=None,
                 return_aux=False):
    """models/blocks.py:237-382 (KPConv.forward).

    q_pts [Nq,3], s_pts [Ns,3], inds [Nq,H] (value Ns = shadow), x [Ns,Cin],
    weights [K,Cin,Cout], kernel_points [K,3].  ``kp_per_query`` [Nq,K,3] replaces the
    shared kernel points (deformable path, blocks.py:286-291) and switches on the
    in-range neighbour filter (blocks.py:300-324).
    """
    Ns = s_pts.shape[0]
    inds = inds.long()
    s_pad = torch.cat([s_pts, torch.full_like(s_pts[:1], SHADOW_COORD)], 0)          # :277
    rel = s_pad[inds] - q_pts[:, None, :]                                             # :280-283  [Nq,H,3]
    kp = kernel_points[None, None] if kp_per_query is None else kp_per_query[:, None]  # [.,1,K,3]
    diff = rel[:, :, None, :] - kp                                                    # :293-294  [Nq,H,K,3]
    sq = (diff ** 2).sum(-1)                                                          # :297      [Nq,H,K]
    aux = {}
    x_pad = torch.cat([x, torch.zeros_like(x[:1])], 0)                                # :356
    if kp_per_query is not None:
        aux["min_d2"] = sq.min(dim=1)[0]                                              # :303
        keep = (sq < extent ** 2).any(dim=2)                                          # :306  [Nq,H]
        # :309-324 re-packs the kept neighbours to the front and points the rest at the
        # shadow row; a shadow neighbour has zero features, so it is equivalent to
        # masking features (and therefore weights*features and the density count).
        nx = x_pad[inds] * keep[:, :, None].to(x.dtype)
    else:
        nx = x_pad[inds]                                                              # :359      [Nq,H,Cin]
    w = kp_influence(sq, extent, influence)                                           # :329-343  [Nq,H,K]
    if aggregation == "closest":                                                      # :346-348
        w = w * F.one_hot(sq.argmin(dim=2), sq.shape[2]).to(w.dtype)
    elif aggregation != "sum":
        raise ValueError(aggregation)
    wf = torch.einsum("nhk,nhc->nkc", w, nx)                                          # :362
    if modulations is not None:
        wf = wf * modulations[:, :, None]                                             # :365-366
    out = torch.einsum("nkc,kco->no", wf, weights)                                    # :369-374
    n = (nx.sum(-1) > 0.0).sum(-1)                                                    # :377-378
    n = torch.clamp(n, min=1).to(out.dtype)                                           # :379
    out = out / n[:, None]                                                            # :380
    if return_aux:
        aux["wf"] = wf
        aux["n"] = n
        return out, aux
    return out


def kpconv(q_pts, s_pts, inds, x, sd, prefix, extent, K=15, influence="linear", aggregation="sum",
           deformable=False, modulated=False, return_aux=False):
    """Full KPConv.forward incl. the offset branch (blocks.py:243-266)."""
    W = sd[prefix + "weights"]
    kp = sd[prefix + "kernel_points"]
    if not deformable:
        return kpconv_rigid(q_pts, s_pts, inds, x, W, kp, extent, influence, aggregation,
                            return_aux=return_aux)
    off_feat = kpconv_rigid(q_pts, s_pts, inds, x, sd[prefix + "offset_conv.weights"],
                            sd[prefix + "offset_conv.kernel_points"], extent, influence,
                            aggregation) + sd[prefix + "offset_bias"]                 # :246
    if modulated:
        unscaled = off_feat[:, :3 * K].reshape(-1, K, 3)                              # :251-252
        mod = 2 * torch.sigmoid(off_feat[:, 3 * K:])                                  # :255
    else:
        unscaled = off_feat.reshape(-1, K, 3)                                         # :260
        mod = None
    deformed = unscaled * extent + kp                                                 # :266, :287
    out = kpconv_rigid(q_pts, s_pts, inds, x, W, kp, extent, influence, aggregation,
                       kp_per_query=deformed, modulations=mod, return_aux=True)
    out[1]["deformed_KP"] = deformed
    out[1]["offset_features"] = off_feat
    return out if return_aux else out[0]


# --------------------------------------------------------------------------- other block ops
def max_pool(x, inds):
    """blocks.py:94-110 (zero shadow row, so an all-negative neighbourhood pools to 0 with shadows)."""
    x_pad = torch.cat([x, torch.zeros_like(x[:1])], 0)
    return x_pad[inds.long()].max(dim=1)[0]


def closest_pool(x, inds):
    """blocks.py:79-91 (first column only)."""
    x_pad = torch.cat([x, torch.zeros_like(x[:1])], 0)
    return x_pad[inds[:, 0].long()]


def unary(x, sd, prefix, relu=True, use_bn=False):
    """blocks.py:481-515 UnaryBlock with use_bn=False: Linear + learned bias (+ LeakyReLU 0.1)."""
    assert not use_bn, "oracle restates the D3Feat default (config.py:42 use_batch_norm=False)"
    y = F.linear(x, sd[prefix + "mlp.weight"], sd[prefix + "mlp.bias"]) + sd[prefix + "batch_norm.bias"]
    return F.leaky_relu(y, 0.1) if relu else y


def block_geometry(batch, layer, strided):
    """blocks.py:588-595 / :660-667."""
    if strided:
        return batch["points"][layer + 1], batch["points"][layer], batch["pools"][layer]
    return batch["points"][layer], batch["points"][layer], batch["neighbors"][layer]


def encoder_plan(config):
    """architectures.py:195-250: (name, layer, radius, in_dim, out_dim) per encoder block + skip indices."""
    plan, skips, skip_dims = [], [], []
    layer, r = 0, config.first_subsampling_dl * config.conv_radius
    in_dim, out_dim = config.in_features_dim, config.first_features_dim
    for bi, name in enumerate(config.architecture):
        if any(t in name for t in ("pool", "strided", "upsample", "global")):
            skips.append(bi)
            skip_dims.append(in_dim)
        if "upsample" in name:
            break
        plan.append((name, layer, r, in_dim, out_dim))
        in_dim = out_dim // 2 if "simple" in name else out_dim
        if "pool" in name or "strided" in name:
            layer += 1
            r *= 2
            out_dim *= 2
    return plan, skips, skip_dims, (layer, r, in_dim, out_dim)


def kpfcnn_forward(sd, batch, config, training=True, collect=None):
    """architectures.py:299-320 KPFCNN.forward -> (features [N,32] L2-normalised, scores [N,1])."""
    plan, skips, skip_dims, (layer, r, in_dim, out_dim) = encoder_plan(config)
    kw = dict(K=config.num_kernel_points, influence=config.KP_influence,
              aggregation=config.aggregation_mode, modulated=config.modulated)
    x = batch["features"].clone().detach()
    skip_x = []
    for bi, (name, lay, rad, din, dout) in enumerate(plan):
        if bi in skips:
            skip_x.append(x)
        pre = "encoder_blocks.%d." % bi
        extent = rad * config.KP_extent / config.conv_radius                          # blocks.py:557,614
        q, s, inds = block_geometry(batch, lay, "strided" in name)
        deform = "deform" in name
        if name.startswith("simple"):                                                 # blocks.py:586-598
            y = kpconv(q, s, inds, x, sd, pre + "KPConv.", extent, deformable=deform, **kw)
            x = F.leaky_relu(y + sd[pre + "batch_norm.bias"], 0.1)
        elif name.startswith("resnetb"):                                              # blocks.py:658-686
            y = unary(x, sd, pre + "unary1.") if din != dout // 4 else x
            y = kpconv(q, s, inds, y, sd, pre + "KPConv.", extent, deformable=deform, **kw)
            if collect is not None:
                collect.append(y)
            y = F.leaky_relu(y + sd[pre + "batch_norm_conv.bias"], 0.1)
            y = unary(y, sd, pre + "unary2.", relu=False)
            sc = max_pool(x, inds) if "strided" in name else x
            if din != dout:
                sc = unary(sc, sd, pre + "unary_shortcut.", relu=False)
            x = F.leaky_relu(y + sc, 0.1)
        elif name in ("max_pool", "max_pool_wide"):
            x = max_pool(x, batch["pools"][lay + 1])
        else:
            raise ValueError("oracle does not restate block " + name)
    # decoder (architectures.py:252-294, :311-314)
    start = next(i for i, b in enumerate(config.architecture) if "upsample" in b)
    for bj, name in enumerate(config.architecture[start:]):
        if bj > 0 and "upsample" in config.architecture[start + bj - 1]:
            x = torch.cat([x, skip_x.pop()], dim=1)
        pre = "decoder_blocks.%d." % bj
        if name == "nearest_upsample":                                                # blocks.py:713
            x = closest_pool(x, batch["upsamples"][layer - 1])
            layer -= 1
        elif name == "unary":
            x = unary(x, sd, pre)
        elif name == "last_unary":                                                    # blocks.py:518-541
            x = F.linear(x, sd[pre + "mlp.weight"], sd[pre + "mlp.bias"])
        else:
            raise ValueError("oracle does not restate block " + name)
    scores = detection_scores(batch["neighbors"][0], x, training)
    return F.normalize(x, p=2, dim=-1), scores


def detection_scores(neighbor, feats, training):
    """architectures.py:322-368."""
    n = feats.shape[0]
    f = torch.cat([feats, torch.zeros_like(feats[:1])], 0)
    nb = torch.cat([neighbor.long(), torch.full_like(neighbor[:1].long(), n)], 0)
    f = f / (f.max() + 1e-6)
    nf = f[nb]                                                                        # [N+1,H,C]
    cnt = (nf.sum(-1) != 0).sum(-1, keepdim=True).clamp(min=1)
    mean = nf.sum(1) / cnt
    local = F.softplus(f - mean)
    depth = f / (1e-6 + f.max(dim=1, keepdim=True)[0])
    scores = (local * depth).max(dim=1, keepdim=True)[0]
    if not training:
        is_max = (f == nf.max(dim=1)[0])
        scores = scores * is_max.float().max(dim=1, keepdim=True)[0]
    return scores[:-1]


# --------------------------------------------------------------------------- losses
def cdist(a, b, metric="euclidean"):
    """utils/loss.py:8-44."""
    if metric == "cosine":
        return torch.sqrt(2 - 2 * a @ b.T)
    if metric == "arccosine":
        return torch.acos(a @ b.T)
    d = a[:, None, :] - b[None, :, :]
    if metric == "sqeuclidean":
        return (d ** 2).sum(-1)
    if metric == "euclidean":
        return torch.sqrt((d ** 2).sum(-1) + 1e-12)
    if metric == "cityblock":
        return d.abs().sum(-1)
    raise NotImplementedError(metric)


def _hardest(d):
    """loss.py:85-87 / :120-121 / :154-155: furthest positive (diagonal), closest negative (off-diagonal)."""
    eye = torch.eye(d.shape[0], dtype=d.dtype, device=d.device)
    return (d * eye).max(dim=1)[0], (d + 1e5 * eye).min(dim=1)[0]


def circle_loss(anchor, positive, dist_keypts, dist_type="euclidean", log_scale=10.0, safe_radius=0.10,
                pos_margin=0.1, neg_margin=1.4):
    """utils/loss.py:111-141.  Returns (loss, accuracy, furthest_positive, average_negative, dists)."""
    d = cdist(anchor, positive, dist_type)
    neg_mask = (dist_keypts > safe_radius)
    fp, cn = _hardest(d)
    avg_neg = (d.sum(-1) - fp) / (d.shape[0] - 1)
    acc = ((fp - cn) < 0).sum() * 100.0 / d.shape[0]
    pos = d - 1e5 * neg_mask.to(d.dtype)
    pw = torch.clamp(pos - pos_margin, min=0).detach()
    zp = log_scale * (pos - pos_margin) * pw
    neg = d + 1e5 * (~neg_mask).to(d.dtype)
    nw = torch.clamp(neg_margin - neg, min=0).detach()
    zn = log_scale * (neg_margin - neg) * nw
    row = F.softplus(torch.logsumexp(zp, -1) + torch.logsumexp(zn, -1)) / log_scale
    col = F.softplus(torch.logsumexp(zp, -2) + torch.logsumexp(zn, -2)) / log_scale
    return (row + col).mean(), acc, fp, avg_neg, d


def contrastive_loss(anchor, positive, dist_keypts, metric="euclidean", pos_margin=0.1, neg_margin=1.4,
                     safe_radius=0.25):
    """utils/loss.py:55-97 (hardest-contrastive).  `dists` returned carries the +10 bumps."""
    d = cdist(anchor, positive, metric)
    P = d.shape[0]
    dk = dist_keypts.detach().to(torch.float64) + 10 * torch.eye(P, dtype=torch.float64)
    d = d + 10 * (dk < safe_radius).to(d.dtype)
    fp, cn = _hardest(d)
    acc = ((fp - cn) < 0).sum() * 100.0 / P
    loss = torch.clamp(fp - pos_margin, min=0) + torch.clamp(neg_margin - cn, min=0)
    avg_neg = (d.sum(-1) - fp) / (P - 1)
    return loss.mean(), acc, fp, avg_neg, d


def det_loss(dists, anc_score, pos_score):
    """utils/loss.py:149-158."""
    fp, cn = _hardest(dists)
    return ((fp - cn) * (anc_score + pos_score).squeeze(-1)).mean()


def pair_losses(features, scores, batch, desc="circle", **kw):
    """trainer.py:90-98 wiring: row-select by corr, descriptor loss, detector loss on its dists."""
    c = batch["corr"].long()
    n0 = int(batch["stack_lengths"][0][0])
    a, p = features[c[:, 0]], features[c[:, 1] + n0]
    sa, sp = scores[c[:, 0]], scores[c[:, 1] + n0]
    fn = circle_loss if desc == "circle" else contrastive_loss
    dl, acc, fp, an, d = fn(a, p, batch["dist_keypts"], **kw)
    return dl, det_loss(d, sa, sp), acc, d


def build_correspondence(source_desc, target_desc):
    """geometric_registration/common.py:5-21 restated (NumPy, fp32 in -> fp32 distances): mutual argmin of
    sqrt(2 - 2 S T^T); np.argmin picks the first NaN of a row/column if there is one."""
    import numpy as np
    s = np.asarray(source_desc)
    t = np.asarray(target_desc)
    with np.errstate(invalid="ignore"):
        distance = np.sqrt(2 - 2 * (s @ t.T))
    source_idx = np.argmin(distance, axis=1)
    target_idx = np.argmin(distance, axis=0)
    result = [[i, source_idx[i]] for i in range(len(source_idx)) if target_idx[source_idx[i]] == i]
    return np.array(result)
