"""Kernel-point dispositions for KPConv.

The reference loads a pre-optimised disposition from kernels/dispositions/k_015_center_3D.ply
and perturbs it at construction (kernels/kernel_points.py:400-482: random z-rotation, N(0, 0.01)
noise, scale by `radius`, drawn from numpy's GLOBAL RNG).  That file is reference data, so this
module computes its own base disposition (centre point + K-1 points spread on a sphere of radius
0.66 by a few hundred steps of Coulomb repulsion, deterministic) and applies the same kind of
perturbation.  ``kernel_points`` is part of the state_dict (blocks.py:234-235), so a reference
checkpoint overrides whatever is generated here.
"""
import functools

import numpy as np

SHELL_RADIUS = 0.66  # |p| of the non-centre points of the reference disposition (unit ball)


@functools.lru_cache(maxsize=None)
def base_disposition(num_kpoints, dimension=3, fixed="center"):
    if dimension != 3:
        raise NotImplementedError("only 3-D kernel dispositions are generated")
    n_fixed = {"center": 1, "verticals": 3, "none": 0}[fixed]
    n_free = num_kpoints - n_fixed
    if n_free < 0:
        raise ValueError("not enough kernel points for fixed=%s" % fixed)
    fixed_pts = np.zeros((n_fixed, 3))
    if fixed == "verticals":
        fixed_pts[1, 2], fixed_pts[2, 2] = SHELL_RADIUS, -SHELL_RADIUS
    # Fibonacci start, then projected gradient descent on sum 1/d over the sphere
    i = np.arange(n_free) + 0.5
    z = 1 - 2 * i / max(n_free, 1)
    phi = i * np.pi * (3 - np.sqrt(5))
    p = np.stack([np.sqrt(1 - z * z) * np.cos(phi), np.sqrt(1 - z * z) * np.sin(phi), z], 1)
    anchors = fixed_pts[1:] / SHELL_RADIUS if fixed == "verticals" else np.zeros((0, 3))
    for it in range(300):
        allp = np.concatenate([p, anchors], 0)
        d = p[:, None, :] - allp[None, :, :]
        r = np.linalg.norm(d, axis=-1) + 1e-9
        np.fill_diagonal(r[:, :n_free], np.inf)
        f = (d / r[..., None] ** 3).sum(1)
        f -= (f * p).sum(1, keepdims=True) * p          # tangential component
        p = p + 0.05 / (1 + it / 50) * f / (np.abs(f).max() + 1e-12)
        p /= np.linalg.norm(p, axis=1, keepdims=True)
    return np.concatenate([fixed_pts, p * SHELL_RADIUS], 0)


def load_kernels(radius, num_kpoints, dimension=3, fixed="center"):
    """Same contract as kernels/kernel_points.py:load_kernels -> float32 [K, dimension]."""
    kp = base_disposition(int(num_kpoints), int(dimension), fixed).copy()
    theta = np.random.rand() * 2 * np.pi
    c, s = np.cos(theta), np.sin(theta)
    rot = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    kp = kp + np.random.normal(scale=0.01, size=kp.shape)
    return (radius * kp @ rot).astype(np.float32)
