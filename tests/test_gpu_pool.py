"""GPU parity of the pooling / gather / detection-score kernels against the torch-CPU oracle (fwd + bwd)."""
import numpy as np
import pytest
import torch

from _util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _case(nq, ns, H, C, seed, dtype=torch.int64):
    rng = np.random.default_rng(seed)
    x = torch.from_numpy(rng.standard_normal((ns, C)).astype(np.float32))
    inds = torch.from_numpy(rng.integers(0, ns + 1, size=(nq, H))).to(dtype)  # ns = shadow
    g = torch.from_numpy(rng.standard_normal((nq, C)).astype(np.float32))
    return x, inds, g


@pytest.mark.parametrize("nq,ns,H,C", [(200, 300, 17, 128), (50, 20, 5, 7), (1000, 4000, 36, 64), (3, 3, 1, 1),
                                       (256, 768, 47, 1024), (100, 150, 70, 260), (33, 40, 32, 130)])
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
def test_max_pool_and_closest_pool(cuda, nq, ns, H, C, dtype):
    from oracle import model_ref
    from d3feat.pytorch_b200.blocks import closest_pool, max_pool
    x, inds, g = _case(nq, ns, H, C, nq + H, dtype)
    for ref_fn, fn in ((model_ref.max_pool, max_pool), (model_ref.closest_pool, closest_pool)):
        xr = x.clone().requires_grad_(True)
        ref = ref_fn(xr, inds)
        (ref * g).sum().backward()
        xg = x.to(cuda).requires_grad_(True)
        out = fn(xg, inds.to(cuda))
        (out * g.to(cuda)).sum().backward()
        assert torch.equal(out.detach().cpu(), ref.detach())          # pure selection: bit-exact
        assert rel_err(xg.grad.cpu(), xr.grad) < TOL


@pytest.mark.parametrize("training", [True, False])
def test_detection_scores(cuda, training):
    from oracle import model_ref
    from d3feat.pytorch_b200 import ops
    rng = np.random.default_rng(11)
    n, H, C = 700, 23, 32
    F = torch.from_numpy(rng.standard_normal((n, C)).astype(np.float32))
    nb = torch.from_numpy(rng.integers(0, n + 1, size=(n, H)))
    nb[:, 0] = torch.arange(n)                     # a point is its own first neighbour, as in the pyramid
    gs = torch.zeros(n, 1)
    sel = rng.choice(n, 64, replace=False)
    gs[sel] = torch.from_numpy(rng.standard_normal((64, 1)).astype(np.float32))
    Fr = F.clone().requires_grad_(True)
    ref = model_ref.detection_scores(nb, Fr, training)
    (ref * gs).sum().backward()
    Fg = F.to(cuda).requires_grad_(True)
    out = ops.detection_scores(Fg, nb.to(cuda), not training)
    (out * gs.to(cuda)).sum().backward()
    assert rel_err(out.detach().cpu(), ref.detach()) < TOL
    assert rel_err(Fg.grad.cpu(), Fr.grad) < TOL


def test_detection_scores_nonpositive_features_use_shadow_max(cuda):
    """All features <= 0: the appended zero shadow row is the global max (architectures.py:330-337)."""
    from oracle import model_ref
    from d3feat.pytorch_b200 import ops
    rng = np.random.default_rng(12)
    F = -torch.from_numpy(rng.random((50, 32)).astype(np.float32)) - 0.1
    nb = torch.from_numpy(rng.integers(0, 51, size=(50, 9)))
    ref = model_ref.detection_scores(nb, F, True)
    out = ops.detection_scores(F.to(cuda), nb.to(cuda), False is True)
    assert rel_err(out.cpu(), ref) < TOL or float((out.cpu() - ref).abs().max()) < 1e-6
