"""Mutual nearest-neighbour matching: oracle vs the reference's outputs (CPU), kernel vs both (GPU)."""
import numpy as np
import pytest
import torch

import _inputs
from conftest import golden


@pytest.mark.parametrize("name", sorted(_inputs.MATCHING_CASES))
def test_oracle_matches_reference_fixture(name):
    from oracle import model_ref
    ns, nt, seed = _inputs.MATCHING_CASES[name]
    s, t = _inputs.matching_case(ns, nt, seed)
    assert np.array_equal(model_ref.build_correspondence(s, t), golden("matching")[name])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(_inputs.MATCHING_CASES))
def test_kernel_matches_reference_fixture(cuda, name):
    """Index work: bit-exact against the reference's output on the seeded sets (the seeds keep nearest and second
    nearest apart by far more than fp32 summation-order noise)."""
    from d3feat.pytorch_b200.matching import build_correspondence
    ns, nt, seed = _inputs.MATCHING_CASES[name]
    s, t = _inputs.matching_case(ns, nt, seed)
    got = build_correspondence(s, t)
    want = golden("matching")[name]
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.gpu
def test_kernel_argmins_nan_rule_and_edges(cuda):
    from oracle import model_ref
    from d3feat.pytorch_b200 import ops
    from d3feat.pytorch_b200.matching import build_correspondence
    rng = np.random.default_rng(0)
    # descriptors a hair longer than 1 and exact duplicates: 2 - 2<s,t> < 0 -> NaN, numpy.argmin takes the first NaN
    s = rng.standard_normal((300, 32)).astype(np.float32)
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    t = s[rng.permutation(300)[:200]].copy() * np.float32(1.0005)
    t[5] = t[7]
    with np.errstate(invalid="ignore"):
        dist = np.sqrt(2 - 2 * (s @ t.T))
    assert np.isnan(dist).any()
    pairs, sarg, targ = ops.mutual_nn(torch.from_numpy(s).to(cuda), torch.from_numpy(t).to(cuda))
    # rows/columns holding a NaN must return its first position; the others the first minimum
    nan_rows = np.isnan(dist).any(1)
    assert np.array_equal(sarg.cpu().numpy()[nan_rows], np.argmin(dist, 1)[nan_rows])
    nan_cols = np.isnan(dist).any(0)
    assert np.array_equal(targ.cpu().numpy()[nan_cols], np.argmin(dist, 0)[nan_cols])
    assert np.array_equal(build_correspondence(s, t), model_ref.build_correspondence(s, t))
    # empty and ragged sizes, non-multiple-of-tile counts, other descriptor widths
    assert build_correspondence(np.zeros((0, 32), np.float32), t).shape == (0,)
    for (a, b, d) in ((1, 1, 32), (129, 257, 32), (513, 100, 16), (77, 300, 64)):
        x, y = _inputs.matching_case(a, b, a + b, dim=d)
        assert np.array_equal(build_correspondence(x, y), model_ref.build_correspondence(x, y)), (a, b, d)
    with pytest.raises(RuntimeError):
        ops.mutual_nn(torch.zeros(3, 32), torch.zeros(3, 32))      # CPU tensors: no CPU path


@pytest.mark.gpu
def test_kernel_full_size_properties(cuda):
    """5000 x 5000 (the reference's evaluation size): every pair is mutual, ascending and unique."""
    from d3feat.pytorch_b200 import ops
    s, t = _inputs.matching_case(5000, 5000, 99)
    pairs, sarg, targ = ops.mutual_nn(torch.from_numpy(s).to(cuda), torch.from_numpy(t).to(cuda))
    p, sa, ta = pairs.cpu().numpy(), sarg.cpu().numpy(), targ.cpu().numpy()
    assert p.shape[0] > 1000
    assert np.all(np.diff(p[:, 0]) > 0) and len(set(p[:, 1].tolist())) == p.shape[0]
    assert np.array_equal(sa[p[:, 0]], p[:, 1]) and np.array_equal(ta[p[:, 1]], p[:, 0])
    mutual = np.nonzero(ta[sa] == np.arange(5000))[0]
    assert np.array_equal(mutual, p[:, 0])
