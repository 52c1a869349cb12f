"""Seeded inputs shared by the golden-fixture generator (oracle/make_golden.py) and the tests,
so fixtures only need to store reference OUTPUTS."""
import numpy as np
import torch


def seeded_state_dict(shapes, seed):
    """shapes: ordered {key: shape}.  Deterministic values independent of torch's RNG/version."""
    rng = np.random.default_rng(seed)
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        if k.endswith("kernel_points"):
            radius = 1.0
            sd[k] = None  # filled by caller (needs the layer radius)
            continue
        fan = shp[-2] if len(shp) >= 2 else max(shp[0], 1)
        if k.endswith("mlp.weight"):
            fan = shp[1]
        bound = 1.0 / np.sqrt(fan)
        if k.endswith("offset_conv.weights"):
            bound *= 0.3
        sd[k] = torch.from_numpy(rng.uniform(-bound, bound, size=shp).astype(np.float32))
    return sd


def unit_kernel_points(K=15):
    """A fixed, well-spread K-point disposition in the unit ball (centre + 2 shells); test data only."""
    pts = [np.zeros(3)]
    n = K - 1
    for i in range(n):
        z = 1 - 2 * (i + 0.5) / n
        r = np.sqrt(max(0.0, 1 - z * z))
        phi = i * np.pi * (3 - np.sqrt(5))
        pts.append(0.66 * np.array([r * np.cos(phi), r * np.sin(phi), z]))
    return np.stack(pts)[:K].astype(np.float64)


def fill_kernel_points(sd, key_radius, K=15, seed=0):
    rng = np.random.default_rng(1000 + seed)
    base = unit_kernel_points(K)
    for k, radius in key_radius.items():
        th = rng.random() * 2 * np.pi
        c, s = np.cos(th), np.sin(th)
        R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        kp = (base + rng.normal(scale=0.01, size=base.shape)) * radius
        sd[k] = torch.from_numpy((kp @ R).astype(np.float32))
    return sd


def kpconv_case(n=2000, cin=64, cout=64, radius=0.075, extent=0.06, seed=0, deformable=False,
                modulated=False, K=15, extent_scale=1.0):
    """BASELINE config 1 style single-layer case: cloud, features, parameters (no neighbours)."""
    rng = np.random.default_rng(seed)
    side = 0.45 * (n / 2000.0) ** (1 / 3)
    pts = (rng.random((n, 3)) * side).astype(np.float32)
    x = rng.standard_normal((n, cin)).astype(np.float32)
    # make a few rows have non-positive channel sums so the density count (blocks.py:377) is exercised
    x[rng.choice(n, n // 10, replace=False)] *= -1.0
    g = rng.standard_normal((n, cout)).astype(np.float32)
    shapes = {"weights": (K, cin, cout)}
    if deformable:
        od = (4 if modulated else 3) * K
        shapes["offset_conv.weights"] = (K, cin, od)
        shapes["offset_bias"] = (od,)
    sd = seeded_state_dict(shapes, seed + 1)
    kr = {"kernel_points": radius}
    if deformable:
        kr["offset_conv.kernel_points"] = radius
        sd["offset_bias"] = torch.from_numpy(rng.uniform(-0.05, 0.05, size=shapes["offset_bias"]).astype(np.float32))
    fill_kernel_points(sd, kr, K, seed)
    return dict(pts=pts, x=x, g=g, sd=sd, radius=radius, extent=extent * extent_scale)


def kpfcnn_shapes(config):
    """All state_dict keys/shapes of the reference KPFCNN (architectures.py:195-294) for `config`
    (use_batch_norm=False), plus {kernel_points key: layer radius}.  Checked against the real
    reference module in oracle/make_golden.py."""
    K = config.num_kernel_points
    shapes, kp_radius = {}, {}
    layer, r = 0, config.first_subsampling_dl * config.conv_radius
    in_dim, out_dim = config.in_features_dim, config.first_features_dim
    skip_dims = []

    def unary(pre, i, o):
        shapes[pre + "mlp.weight"] = (o, i)
        shapes[pre + "mlp.bias"] = (o,)
        shapes[pre + "batch_norm.bias"] = (o,)

    def conv(pre, i, o, deform):
        shapes[pre + "weights"] = (K, i, o)
        kp_radius[pre + "kernel_points"] = r
        if deform:
            od = (4 if config.modulated else 3) * K
            shapes[pre + "offset_conv.weights"] = (K, i, od)
            kp_radius[pre + "offset_conv.kernel_points"] = r
            shapes[pre + "offset_bias"] = (od,)

    nblk = 0
    for bi, name in enumerate(config.architecture):
        if any(t in name for t in ("pool", "strided", "upsample", "global")):
            skip_dims.append(in_dim)
        if "upsample" in name:
            break
        pre = "encoder_blocks.%d." % bi
        deform = "deform" in name
        if name.startswith("simple"):
            conv(pre + "KPConv.", in_dim, out_dim // 2, deform)
            shapes[pre + "batch_norm.bias"] = (out_dim // 2,)
        elif name.startswith("resnetb"):
            if in_dim != out_dim // 4:
                unary(pre + "unary1.", in_dim, out_dim // 4)
            conv(pre + "KPConv.", out_dim // 4, out_dim // 4, deform)
            shapes[pre + "batch_norm_conv.bias"] = (out_dim // 4,)
            unary(pre + "unary2.", out_dim // 4, out_dim)
            if in_dim != out_dim:
                unary(pre + "unary_shortcut.", in_dim, out_dim)
        in_dim = out_dim // 2 if "simple" in name else out_dim
        if "pool" in name or "strided" in name:
            layer += 1
            r *= 2
            out_dim *= 2
        nblk += 1
    start = next(i for i, b in enumerate(config.architecture) if "upsample" in b)
    for bj, name in enumerate(config.architecture[start:]):
        if bj > 0 and "upsample" in config.architecture[start + bj - 1]:
            in_dim += skip_dims[layer]
        pre = "decoder_blocks.%d." % bj
        if name == "unary":
            unary(pre, in_dim, out_dim)
        elif name == "last_unary":
            shapes[pre + "mlp.weight"] = (32, in_dim)
            shapes[pre + "mlp.bias"] = (32,)
        in_dim = out_dim
        if "upsample" in name:
            layer -= 1
            r *= 0.5
            out_dim = out_dim // 2
    return shapes, kp_radius


def kpfcnn_state_dict(config, seed=0):
    shapes, kp_radius = kpfcnn_shapes(config)
    sd = seeded_state_dict(shapes, seed)
    rng = np.random.default_rng(77 + seed)
    for k in sorted(shapes):
        if k.endswith("bias"):
            sd[k] = torch.from_numpy(rng.uniform(-0.05, 0.05, size=shapes[k]).astype(np.float32))
    fill_kernel_points(sd, kp_radius, config.num_kernel_points, seed)
    return sd


# ------------------------------------------------------------------------------------------- descriptor matching (f3)
MATCHING_CASES = {"m_700_650": (700, 650, 11), "m_250_250": (250, 250, 12), "m_1_40": (1, 40, 13), "m_2000_1800": (2000, 1800, 14)}


def matching_case(ns, nt, seed, dim=32, overlap=0.6, noise=0.15):
    """Two sets of unit descriptors sharing `overlap` of their points up to noise (so mutual matches exist)."""
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((max(ns, nt), dim))
    s = base[:ns] + 0.0 * rng.standard_normal((ns, dim))
    t = base[:nt].copy()
    fresh = rng.random(nt) > overlap
    t[fresh] = rng.standard_normal((int(fresh.sum()), dim))
    t += noise * rng.standard_normal((nt, dim))
    t = t[rng.permutation(nt)]
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    return s.astype(np.float32), t.astype(np.float32)
