"""CPU suite (no GPU): the oracle against the committed reference fixtures, the host logic, and the
C-ABI library's exported symbols.  No compute call touches the CUDA library here."""
import os
import re

import numpy as np
import pytest
import torch

import _inputs
from conftest import ROOT, golden
from _util import canonical_rows, rel_err
from d3feat.pytorch_b200 import synthetic
from d3feat.pytorch_b200.config import build_architecture, default_config
from oracle import model_ref


def _two_fragments(n0, n1, seed):
    p = np.concatenate([synthetic.room_shell_fragment(n0, seed), synthetic.room_shell_fragment(n1, seed + 1)])
    return p, np.array([n0, n1], np.int32)


def test_oracle_port_matches_reference_fixture(oracle_cpu):
    """oracle/d3feat_oracle.c vs fixtures written by the unmodified reference C++."""
    g = golden("native_pyramid")
    p, lens = _two_fragments(1800, 1400, 11)
    r = 0.075
    for lvl in range(3):
        nb = oracle_cpu.batch_query(p, p, lens, lens, r)
        assert np.array_equal(nb, canonical_rows(p, p, g["nb%d" % lvl])[0])
        sp, sl = oracle_cpu.subsample_batch(p, lens, 2 * r / 2.5)
        assert np.array_equal(sl, g["sublen%d" % lvl])
        assert np.array_equal(sp.view(np.uint32), g["sub%d" % lvl].view(np.uint32))
        assert np.array_equal(oracle_cpu.batch_query(sp, p, sl, lens, r), canonical_rows(sp, p, g["pool%d" % lvl])[0])
        assert np.array_equal(oracle_cpu.batch_query(p, sp, lens, sl, 2 * r), canonical_rows(p, sp, g["up%d" % lvl])[0])
        p, lens, r = sp, sl, 2 * r


def test_oracle_port_matches_live_reference_library(oracle_cpu):
    """Where oracle/_ref (the reference C++ itself) is present, compare live on fresh inputs."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libd3feat_ref.so")):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(123)
    for n, dl in [(60, 0.08), (5000, 0.05), (30000, 0.03)]:
        p = (rng.random((n, 3)) * 1.2 - 0.4).astype(np.float32)
        lens = np.array([n // 3, n - n // 3], np.int32)
        a, al = oracle_cpu.subsample_batch(p, lens, dl, impl="ref")
        b, bl = oracle_cpu.subsample_batch(p, lens, dl, impl="port")
        assert np.array_equal(al, bl) and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    p = (rng.random((5000, 3)) * 0.7).astype(np.float32)
    lens = np.array([2000, 3000], np.int32)
    ref = oracle_cpu.batch_query(p, p, lens, lens, 0.075, impl="ref")
    port = oracle_cpu.batch_query(p, p, lens, lens, 0.075, impl="port")
    assert ref.shape == port.shape and np.array_equal(canonical_rows(p, p, ref)[0], port)


def test_unordered_map_order_small_and_rehash_boundaries(oracle_cpu):
    """Cell counts around every rehash threshold of libstdc++'s prime policy (13, 29, 59, ...)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libd3feat_ref.so")):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    for m in [1, 2, 12, 13, 14, 28, 29, 30, 58, 59, 60, 126, 127, 128, 256, 257, 258, 540, 541, 542, 1110]:
        cells = rng.choice(40 ** 3, m, replace=False)
        c = np.stack([cells % 40, (cells // 40) % 40, cells // 1600], 1)
        p = ((c + rng.random((m, 3)) * 0.9 + 0.05) * 0.1).astype(np.float32)
        p = np.concatenate([p, p[: m // 2] + np.float32(0.001)])
        p = p[rng.permutation(p.shape[0])]
        lens = np.array([p.shape[0]], np.int32)
        a, al = oracle_cpu.subsample_batch(p, lens, 0.1, impl="ref")
        b, bl = oracle_cpu.subsample_batch(p, lens, 0.1, impl="port")
        assert np.array_equal(al, bl) and np.array_equal(a.view(np.uint32), b.view(np.uint32)), m


KP_CASES = [
    ("kpconv_rigid_2k", 2000, 64, 64, False, False, "linear", "sum", 100),
    ("kpconv_rigid_c1", 1500, 1, 64, False, False, "linear", "sum", 101),
    ("kpconv_gauss_closest", 600, 16, 24, False, False, "gaussian", "closest", 102),
    ("kpconv_constant", 600, 16, 8, False, False, "constant", "sum", 103),
    ("kpconv_deform", 800, 32, 32, True, False, "linear", "sum", 104),
    ("kpconv_deform_mod", 800, 16, 32, True, True, "linear", "sum", 105),
    ("kpconv_deform_gauss", 500, 16, 16, True, False, "gaussian", "sum", 106),
]


@pytest.mark.parametrize("name,n,cin,cout,deform,mod,infl,agg,seed", KP_CASES)
def test_oracle_kpconv_matches_reference_fixture(name, n, cin, cout, deform, mod, infl, agg, seed):
    g = golden(name)
    case = _inputs.kpconv_case(n=n, cin=cin, cout=cout, seed=seed, deformable=deform, modulated=mod)
    if cin == 1:
        case["x"] = np.ones_like(case["x"])
    sd = {k: v.clone().requires_grad_("kernel_points" not in k) for k, v in case["sd"].items()}
    pts = torch.from_numpy(case["pts"])
    x = torch.from_numpy(case["x"]).requires_grad_(True)
    out = model_ref.kpconv(pts, pts, torch.from_numpy(g["inds"]), x, sd, "", case["extent"], influence=infl,
                           aggregation=agg, deformable=deform, modulated=mod, return_aux=deform)
    aux = None
    if deform:
        out, aux = out
    (out * torch.from_numpy(case["g"])).sum().backward()
    assert rel_err(out.detach(), g["out"]) < 2e-5
    assert rel_err(x.grad, g["dx"]) < 2e-5
    assert rel_err(sd["weights"].grad, g["d_weights"]) < 2e-5
    if deform:
        assert rel_err(sd["offset_conv.weights"].grad, g["d_offset_conv__weights"]) < 2e-5
        assert rel_err(aux["min_d2"].detach(), g["min_d2"]) < 2e-5


def test_oracle_losses_match_reference_fixture():
    g = golden("losses")
    rng = np.random.default_rng(42)
    for P in (128, 64, 7):
        a = rng.standard_normal((P, 32)); a /= np.linalg.norm(a, axis=1, keepdims=True)
        p = a + 0.25 * rng.standard_normal((P, 32)); p /= np.linalg.norm(p, axis=1, keepdims=True)
        kp = rng.random((P, 3)) * 0.6
        dk = torch.from_numpy(np.sqrt(((kp[:, None] - kp[None]) ** 2).sum(-1)))
        sa, sp = rng.random((P, 1)).astype(np.float32), rng.random((P, 1)).astype(np.float32)
        for kind, fn in (("circle", model_ref.circle_loss), ("contrastive", model_ref.contrastive_loss)):
            A = torch.from_numpy(a.astype(np.float32)).requires_grad_(True)
            B = torch.from_numpy(p.astype(np.float32)).requires_grad_(True)
            l, acc, fp, an, d = fn(A, B, dk, safe_radius=0.1)
            det = model_ref.det_loss(d, torch.from_numpy(sa), torch.from_numpy(sp))
            (l + det).backward()
            pre = "%s%d_" % (kind, P)
            assert rel_err(l.detach(), g[pre + "loss"]) < 1e-5 and rel_err(det.detach(), g[pre + "det"]) < 1e-5
            assert rel_err(A.grad, g[pre + "dA"]) < 1e-5 and rel_err(B.grad, g[pre + "dB"]) < 1e-5


def test_oracle_kpfcnn_matches_reference_fixture(oracle_cpu):
    """Whole pair through the oracle: C pyramid + torch-CPU KPFCNN + losses vs the reference's outputs."""
    g = golden("kpfcnn_rigid")
    cfg = default_config(first_features_dim=32)
    data = synthetic.fragment_pair(1500, seed=5, num_node=64)
    from oracle.pipeline import cpu_collate
    batch = cpu_collate(data, cfg, [40] * 5, impl="port")
    assert [p.shape[0] for p in batch["points"]] == g["N"].tolist()
    assert [int(x.sum()) for x in batch["neighbors"]] == g["nb_sum"].tolist()
    sd = _inputs.kpfcnn_state_dict(cfg, seed=3)
    with torch.no_grad():
        f, s = model_ref.kpfcnn_forward(sd, batch, cfg, training=True)
        dl, det, acc, _ = model_ref.pair_losses(f, s, batch, "circle")
    assert rel_err(f, g["features"]) < 2e-5 and rel_err(s, g["scores"]) < 2e-5
    assert rel_err(dl, g["desc_loss"]) < 2e-5 and rel_err(det, g["det_loss"]) < 2e-5


# ----------------------------------------------------------------------------- host logic
def test_architecture_builder_and_state_dict_keys():
    from d3feat.pytorch_b200.architectures import KPFCNN
    arch = build_architecture(5)
    assert arch[:3] == ["simple", "resnetb", "resnetb_strided"] and arch[-1] == "last_unary" and len(arch) == 22
    assert build_architecture(5, deformable_from=3).count("resnetb_deformable") == 4
    for kw in ({}, dict(architecture=build_architecture(5, deformable_from=3)), dict(modulated=True, architecture=build_architecture(5, 3))):
        cfg = default_config(first_features_dim=32, **kw)
        net = KPFCNN(cfg)
        shapes, kpr = _inputs.kpfcnn_shapes(cfg)
        sd = net.state_dict()
        assert set(sd) == set(shapes) | set(kpr)
        assert all(tuple(sd[k].shape) == tuple(v) for k, v in shapes.items())
    full = KPFCNN(default_config())
    assert sum(p.numel() for p in full.parameters() if p.requires_grad) == 24316320


def test_kpconv_module_surface():
    from d3feat.pytorch_b200.blocks import KPConv, block_decider
    np.random.seed(0)
    m = KPConv(15, 3, 8, 16, 0.06, 0.075, deformable=True, modulated=True)
    assert set(m.state_dict()) == {"weights", "kernel_points", "offset_bias", "offset_conv.weights", "offset_conv.kernel_points"}
    assert m.offset_dim == 60 and m.kernel_points.shape == (15, 3) and not m.kernel_points.requires_grad
    assert abs(float(m.kernel_points.norm(dim=1).max()) - 0.66 * 0.075) < 0.01
    assert repr(m) == "KPConv(radius: 0.07, extent: 0.06, in_feat: 8, out_feat: 16)" or "KPConv(radius" in repr(m)
    with pytest.raises(ValueError):
        block_decider("nonsense", 0.1, 4, 4, 0, default_config())
    with pytest.raises(ValueError):
        KPConv(15, 3, 4, 4, 0.06, 0.075, KP_influence="cubic")


def test_synthetic_generator_is_deterministic():
    a = synthetic.fragment_pair(800, seed=3)
    b = synthetic.fragment_pair(800, seed=3)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert a[0].shape == (800, 3) and a[0].dtype == np.float32 and a[4].shape == (128, 2) and a[5].dtype == np.float64


def test_c_abi_library_exports_every_declared_symbol(built_lib):
    """include/d3feat_b200.h <-> libd3feat_b200.so <-> the ctypes table; no compute call is made."""
    from d3feat.pytorch_b200 import _lib
    header = open(os.path.join(ROOT, "include", "d3feat_b200.h")).read()
    debug = open(os.path.join(ROOT, "include", "d3feat_b200_debug.h")).read()
    product = set(re.findall(r"\b(d3f_[a-z0-9_]+)\s*\(", header))
    declared = product | set(re.findall(r"\b(d3f_[a-z0-9_]+)\s*\(", debug))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    # the product ABI holds no process-global selector: those live in the debug header only
    assert not [n for n in product if n.startswith("d3f_set_") or n.endswith("_set_gather_events")]
    for name in declared:
        assert hasattr(built_lib, name), name
    assert built_lib.d3f_version() == 100
    assert built_lib.d3f_pair_loss_aux_floats(128) >= 128 * 128
    assert built_lib.d3f_radius_neighbors_workspace_bytes(1000, 1000, 2) > 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "d3feat")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)
