"""Host-side logic of the static engine and the bench helpers (no GPU): pyramid plan, capacities, byte accounting."""
import importlib.util
import os

from d3feat.pytorch_b200.config import default_config
from d3feat.pytorch_b200.engine import _pyramid_levels, plan_capacities

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pyramid_levels_follow_the_reference_block_walk():
    """dataloader.py:95-170: one level per pooling/strided block plus the last encoder level; default arch = 5 levels,
    4 of them pooled, none deformable."""
    lv = _pyramid_levels(default_config())
    assert len(lv) == 5
    assert [l[0] for l in lv] == [True] * 5                 # every level has convolutions
    assert [l[2] for l in lv] == [True, True, True, True, False]
    assert not any(l[1] or l[3] for l in lv)
    arch = ['simple', 'resnetb', 'resnetb_strided', 'resnetb_deformable', 'resnetb', 'resnetb_deformable_strided',
            'resnetb_deformable', 'nearest_upsample', 'unary', 'nearest_upsample', 'unary', 'last_unary']
    lv = _pyramid_levels(default_config(architecture=arch))
    assert len(lv) == 3
    # a level's conv search uses the deformable radius if a block BEFORE its last one is deformable
    # (`layer_blocks[:-1]`, dataloader.py:103-108); the pool search if the strided block itself is deformable
    assert [l[1] for l in lv] == [False, True, False]
    assert [l[3] for l in lv] == [False, True, False]


def test_plan_capacities_margin_and_alignment():
    caps = plan_capacities([[40000, 12065, 2515, 693, 190], [40000, 12094, 2508, 689, 187]], margin=1.10, align=64)
    assert caps[0] == 40000
    for c, m in zip(caps[1:], (12094, 2515, 693, 190)):
        assert c % 64 == 0 and c >= 1.10 * m and c < 1.10 * m + 64


def test_bench_byte_accounting_matches_survey():
    """L0 resnetb 32->32 at Nq = Ns = 40000, H = 36 is 213 MB; the bench uses the same formula."""
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    src = open(os.path.join(ROOT, "bench.py")).read()
    ns = {}
    start = src.index("def kpconv_logical_bytes")
    end = src.index("def make_pairs")
    exec(src[start:end], ns)
    assert abs(ns["kpconv_logical_bytes"](40000, 40000, 36, 32, 32) / 1e6 - 213.0) < 0.6
    assert ns["kpconv_logical_bytes"](40000, 40000, 35, 32, 32) == 207261440
    assert abs(ns["kpconv_flops"](40000, 36, 32, 32) / 1e9 - 2.87) < 0.05
    assert spec is not None


def test_zero_arena_sizes_itself_from_the_previous_step_and_hands_out_cleared_views():
    """ops.ZeroArena (one fill per step for every split-K GEMM output): the first step spills to torch.zeros, reset() then
    allocates what that step needed, later steps take 128-byte-aligned views of the cleared buffer, a request that does
    not fit still gets zeros, and reset() clears exactly what was handed out."""
    import torch
    from d3feat.pytorch_b200.ops import ZeroArena
    dev = torch.device("cpu")
    ar = ZeroArena(dev)
    ar.reset()
    a = ar.take(100, dev); b = ar.take(7, dev)
    assert ar.buf is None and ar.spilled == 128 + 32 and float(a.abs().sum() + b.abs().sum()) == 0.0
    ar.reset()                                   # sized by the step before: 160 floats
    assert ar.buf.numel() == 160 and ar.off == 0
    a = ar.take(100, dev); b = ar.take(7, dev)
    assert a.data_ptr() == ar.buf.data_ptr() and b.data_ptr() == ar.buf.data_ptr() + 128 * 4 and ar.spilled == 0
    a.fill_(3.0); b.fill_(5.0)
    c = ar.take(50, dev)                         # does not fit any more: falls back to fresh zeros
    assert ar.spilled == 64 and c.data_ptr() != ar.buf.data_ptr() and float(c.abs().sum()) == 0.0
    ar.reset()                                   # grows to 224 floats (fresh zeros)
    assert ar.buf.numel() == 224 and float(ar.buf.abs().sum()) == 0.0
    x = ar.take(224, dev); x.fill_(1.0)
    ar.reset()                                   # same size: cleared in place
    assert ar.buf.numel() == 224 and float(ar.buf.abs().sum()) == 0.0 and ar.high == 224
