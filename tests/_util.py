import numpy as np


def d2_rows(q, s, idx):
    """fp32 squared distances of neighbour matrix rows, reference arithmetic (no FMA); pad -> inf."""
    ns = s.shape[0]
    sp = np.concatenate([s, np.full((1, 3), np.inf, np.float32)], 0)
    nbr = sp[np.minimum(idx, ns)]
    d = (q[:, None, :] - nbr).astype(np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    d2 = np.where(idx >= ns, np.inf, d2).astype(np.float32)
    return d2


def canonical_rows(q, s, idx):
    """Order every row by (d2, index): the documented tie contract (ties are the only place the
    reference's unstable std::sort can differ)."""
    d2 = d2_rows(q, s, idx)
    order = np.lexsort((idx, d2))  # sorts along the last axis, primary key d2, secondary idx
    return np.take_along_axis(idx, order, axis=1), d2


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
