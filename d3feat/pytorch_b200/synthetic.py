"""Deterministic synthetic 'room-shell' fragments.

The 3DMatch data the reference trains on (datasets/ThreeDMatch.py:93-149) is not
available offline; this generator produces the same *tuple contract* the
reference dataset hands to ``collate_fn_descriptor``
(pts0, pts1, feat0, feat1, sel_corr, dist_keypts) from seeded numpy only.
"""
import numpy as np


def room_shell_fragment(n, seed, voxel=0.03, noise=0.004):
    """n points on the faces of an axis-aligned box, voxel-thinned to `voxel`, fp32 [n,3].

    Mimics a 3DMatch fragment after the 0.03 m voxel down-sampling the reference
    applies (config.py:79 `downsample`, ThreeDMatch.py:191).
    """
    rng = np.random.default_rng(seed)
    # 20k points <-> a 1.5 x 1.5 x 1.25 m room; other sizes keep the same surface density
    room = np.array([1.5, 1.5, 1.25])
    scale = float(np.sqrt(n / 20000.0))
    while True:
        ext = room * scale
        m = 8 * n
        face = rng.integers(0, 6, size=m)
        p = rng.random((m, 3)) * ext
        axis = face % 3
        side = (face // 3).astype(np.float64)
        p[np.arange(m), axis] = side * ext[axis]
        p += rng.normal(0.0, noise, size=p.shape)
        p = p.astype(np.float32)
        # voxel thinning: keep the first point of every occupied voxel
        key = np.floor(p / np.float32(voxel)).astype(np.int64)
        key -= key.min(axis=0)
        dims = key.max(axis=0) + 1
        lin = (key[:, 0] * dims[1] + key[:, 1]) * dims[2] + key[:, 2]
        _, first = np.unique(lin, return_index=True)
        first.sort()
        p = p[first]
        if p.shape[0] >= n:
            break
        scale *= 1.15  # room too small for n voxels: enlarge and redraw
    p = p[rng.permutation(p.shape[0])[:n]]
    return np.ascontiguousarray(p, dtype=np.float32)


def fragment_pair(n, seed, num_node=128):
    """The tuple ThreeDMatchDataset.__getitem__ returns (ThreeDMatch.py:147-149), synthetic."""
    pts0 = room_shell_fragment(n, 2 * seed)
    pts1 = room_shell_fragment(n, 2 * seed + 1)
    rng = np.random.default_rng(10_000 + seed)
    feat0 = np.ones((n, 1), np.float32)   # ThreeDMatch.py:143-144
    feat1 = np.ones((n, 1), np.float32)
    P = min(num_node, n)
    sel0 = rng.choice(n, P, replace=False)
    sel1 = rng.choice(n, P, replace=False)
    sel_corr = np.stack([sel0, sel1], axis=1).astype(np.int64)
    a = pts0[sel0].astype(np.float64)
    d = a[:, None, :] - a[None, :, :]
    dist_keypts = np.sqrt((d * d).sum(-1))  # scipy cdist of the source keypoints (ThreeDMatch.py:137)
    return pts0, pts1, feat0, feat1, sel_corr, dist_keypts


def random_cloud(n, seed, extent=0.45):
    """BASELINE config 1: uniform random cloud in [0, extent]^3."""
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3)) * extent).astype(np.float32)
