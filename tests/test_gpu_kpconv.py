"""GPU parity of the KPConv kernels: reference fixtures (fwd + bwd) and random shapes vs the CPU oracle."""
import numpy as np
import pytest
import torch

import _inputs
from conftest import golden
from _util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4  # BASELINE.json north_star: descriptors and loss within 1e-4 relative, fp32


@pytest.fixture(params=[1, 2, 3], ids=["ffma+gemm", "mma+gemm", "fused"])
def kp_impl(request, built_lib):
    """Every KPConv parity case runs on every forward path (debug selector d3f_set_kpconv_impl): gather kernel with FFMA or
    mma.sync correlation + contraction GEMM, and the fused kernel (falls back to mma+gemm where a layer is not eligible)."""
    built_lib.d3f_set_kpconv_impl(request.param)
    assert built_lib.d3f_get_kpconv_impl() == request.param
    yield request.param
    built_lib.d3f_set_kpconv_impl(-1)

CASES = [
    ("kpconv_rigid_2k", 2000, 64, 64, False, False, "linear", "sum", 100),
    ("kpconv_rigid_c1", 1500, 1, 64, False, False, "linear", "sum", 101),
    ("kpconv_gauss_closest", 600, 16, 24, False, False, "gaussian", "closest", 102),
    ("kpconv_constant", 600, 16, 8, False, False, "constant", "sum", 103),
    ("kpconv_deform", 800, 32, 32, True, False, "linear", "sum", 104),
    ("kpconv_deform_mod", 800, 16, 32, True, True, "linear", "sum", 105),
    ("kpconv_deform_gauss", 500, 16, 16, True, False, "gaussian", "sum", 106),
]


@pytest.mark.parametrize("name,n,cin,cout,deform,mod,infl,agg,seed", CASES)
@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
def test_kpconv_module_vs_reference_fixture(cuda, kp_impl, name, n, cin, cout, deform, mod, infl, agg, seed, idx_dtype):
    from d3feat.pytorch_b200.blocks import KPConv
    g = golden(name)
    case = _inputs.kpconv_case(n=n, cin=cin, cout=cout, seed=seed, deformable=deform, modulated=mod)
    if cin == 1:
        case["x"] = np.ones_like(case["x"])
    m = KPConv(15, 3, cin, cout, case["extent"], case["radius"], KP_influence=infl, aggregation_mode=agg,
               deformable=deform, modulated=mod).to(cuda)
    m.load_state_dict(case["sd"], strict=True)
    pts = torch.from_numpy(case["pts"]).to(cuda)
    x = torch.from_numpy(case["x"]).to(cuda).requires_grad_(True)
    inds = torch.from_numpy(g["inds"]).to(cuda).to(idx_dtype)
    out = m(pts, pts, inds, x)
    (out * torch.from_numpy(case["g"]).to(cuda)).sum().backward()
    errs = {"out": rel_err(out.detach().cpu(), g["out"]), "dx": rel_err(x.grad.cpu(), g["dx"])}
    for k, p in m.named_parameters():
        if p.grad is not None:
            errs["d_" + k] = rel_err(p.grad.cpu(), g["d_" + k.replace(".", "__")])
    if deform:
        errs["min_d2"] = rel_err(m.min_d2.cpu(), g["min_d2"])
        errs["deformed_KP"] = rel_err(m.deformed_KP.detach().cpu(), g["deformed_KP"])
    print(name, {k: "%.1e" % v for k, v in errs.items()})
    assert max(errs.values()) < TOL, errs


@pytest.mark.parametrize("nq,ns,H,cin,cout", [(1, 1, 1, 1, 1), (37, 50, 7, 3, 45), (300, 200, 40, 33, 64),
                                              (129, 257, 19, 130, 17), (64, 64, 48, 256, 128), (5, 9, 0, 8, 8),
                                              (200, 300, 35, 32, 32), (90, 120, 42, 64, 64), (70, 80, 9, 128, 16),
                                              (33, 40, 70, 12, 20), (10, 1, 5, 8, 8)])
def test_kpconv_random_shapes_vs_oracle(cuda, kp_impl, nq, ns, H, cin, cout):
    """Ragged shapes, shadow indices, strided index rows; CUDA vs the torch-CPU restatement."""
    from oracle import model_ref
    from d3feat.pytorch_b200.blocks import _KPConvFunction
    rng = np.random.default_rng(nq * 1000 + H)
    q = torch.from_numpy((rng.random((nq, 3)) * 0.2).astype(np.float32))
    s = torch.from_numpy((rng.random((ns, 3)) * 0.2).astype(np.float32))
    wide = torch.from_numpy(rng.integers(0, ns + 1, size=(nq, H + 3)).astype(np.int64))  # ns = shadow
    inds = wide[:, :H]                                                                   # non-contiguous rows
    x = torch.from_numpy(rng.standard_normal((ns, cin)).astype(np.float32)).requires_grad_(True)
    W = torch.from_numpy((rng.standard_normal((15, cin, cout)) / np.sqrt(15 * cin)).astype(np.float32)).requires_grad_(True)
    kp = torch.from_numpy((_inputs.unit_kernel_points() * 0.075).astype(np.float32))
    gout = torch.from_numpy(rng.standard_normal((nq, cout)).astype(np.float32))
    ref = model_ref.kpconv_rigid(q, s, inds, x, W, kp, 0.06)
    (ref * gout).sum().backward()
    xg = x.detach().to(cuda).requires_grad_(True)
    Wg = W.detach().to(cuda).requires_grad_(True)
    out, _ = _KPConvFunction.apply(q.to(cuda), s.to(cuda), wide.to(cuda)[:, :H], xg, Wg, kp.to(cuda), None, 0.06,
                                   "linear", "sum", False, False)
    (out * gout.to(cuda)).sum().backward()
    scale = max(float(ref.detach().abs().max()), 1e-6)
    assert float((out.cpu() - ref.detach()).abs().max()) / scale < TOL
    assert rel_err(xg.grad.cpu(), x.grad) < TOL if x.grad.abs().max() > 0 else float(xg.grad.abs().max()) == 0
    assert rel_err(Wg.grad.cpu(), W.grad) < TOL if W.grad.abs().max() > 0 else float(Wg.grad.abs().max()) == 0


def test_kpconv_rejects_cpu_tensors(built_lib):
    from d3feat.pytorch_b200.blocks import KPConv
    m = KPConv(15, 3, 4, 4, 0.06, 0.075)
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        m(torch.zeros(3, 3), torch.zeros(3, 3), torch.zeros(3, 2, dtype=torch.long), torch.zeros(3, 4))


@pytest.mark.parametrize("deform", [False, True])
def test_kpconv_fused_bias_activation_matches_unfused(cuda, kp_impl, deform):
    """KPConv(..., bias=b, slope=0.1) (epilogue of the contraction) == leaky_relu(KPConv(...) + b), values and grads."""
    from d3feat.pytorch_b200.blocks import KPConv
    g = golden("kpconv_deform" if deform else "kpconv_rigid_2k")
    n, c = (800, 32) if deform else (2000, 64)
    case = _inputs.kpconv_case(n=n, cin=c, cout=c, seed=104 if deform else 100, deformable=deform, modulated=False)
    m = KPConv(15, 3, c, c, case["extent"], case["radius"], deformable=deform).to(cuda)
    m.load_state_dict(case["sd"], strict=True)
    pts = torch.from_numpy(case["pts"]).to(cuda)
    inds = torch.from_numpy(g["inds"]).to(cuda)
    gout = torch.from_numpy(case["g"]).to(cuda)
    res = []
    for fused in (False, True):
        m.zero_grad(set_to_none=True)
        x = torch.from_numpy(case["x"]).to(cuda).requires_grad_(True)
        b = torch.linspace(-0.05, 0.05, c, device=cuda).requires_grad_(True)
        if fused:
            out = m(pts, pts, inds, x, bias=b, slope=0.1)
        else:
            out = torch.nn.functional.leaky_relu(m(pts, pts, inds, x) + b, 0.1)
        (out * gout).sum().backward()
        res.append((out.detach(), x.grad, b.grad, m.weights.grad.clone()))
    for a, r in zip(res[1], res[0]):
        assert rel_err(a.cpu(), r.cpu()) < 1e-5


@pytest.mark.parametrize("nq,ns,H,cin,cout", [(200, 300, 35, 32, 32), (90, 120, 42, 64, 64), (64, 64, 48, 256, 128),
                                              (300, 200, 40, 33, 64), (500, 100, 70, 16, 32), (40, 1, 5, 8, 32),
                                              (50, 60, 9, 12, 256)])
@pytest.mark.parametrize("idx_dtype", [torch.int64, torch.int32])
def test_kpconv_backward_over_transposed_lists(cuda, nq, ns, H, cin, cout, idx_dtype):
    """Atomic-free grad_x (gather over d3f_neighbors_transpose lists + GEMM with W^T) vs the CPU oracle and vs the
    scatter path; in-degrees far above one 64-entry chunk are covered by (500, 100, 70) and (40, 1, 5)."""
    from oracle import model_ref
    from d3feat.pytorch_b200 import ops
    from d3feat.pytorch_b200.blocks import _KPConvFunction
    rng = np.random.default_rng(nq * 31 + H)
    q = torch.from_numpy((rng.random((nq, 3)) * 0.2).astype(np.float32))
    s = torch.from_numpy((rng.random((ns, 3)) * 0.2).astype(np.float32))
    inds = torch.from_numpy(rng.integers(0, ns + 1, size=(nq, H)).astype(np.int64))      # ns = shadow
    x = torch.from_numpy(rng.standard_normal((ns, cin)).astype(np.float32)).requires_grad_(True)
    W = torch.from_numpy((rng.standard_normal((15, cin, cout)) / np.sqrt(15 * cin)).astype(np.float32)).requires_grad_(True)
    kp = torch.from_numpy((_inputs.unit_kernel_points() * 0.075).astype(np.float32))
    gout = torch.from_numpy(rng.standard_normal((nq, cout)).astype(np.float32))
    ref = model_ref.kpconv_rigid(q, s, inds, x, W, kp, 0.06)
    (ref * gout).sum().backward()
    inds_g = inds.to(cuda).to(idx_dtype)
    t_off, t_src = ops.neighbors_transpose(inds_g, ns)
    # the lists hold every (query, support) reference exactly once
    off = t_off.cpu().numpy()
    assert off[0] == 0 and off[-1] == int((inds < ns).sum())
    counts = np.bincount(inds.numpy()[inds.numpy() < ns].ravel(), minlength=ns)
    assert np.array_equal(np.diff(off), counts)
    src = t_src.cpu().numpy()
    for j in range(0, ns, max(1, ns // 7)):
        assert sorted(src[off[j]:off[j + 1]].tolist()) == sorted(np.nonzero((inds.numpy() == j))[0].tolist())
    grads = []
    for tr in ((None, None), (t_off, t_src)):
        xg = x.detach().to(cuda).requires_grad_(True)
        Wg = W.detach().to(cuda).requires_grad_(True)
        out, _ = _KPConvFunction.apply(q.to(cuda), s.to(cuda), inds_g, xg, Wg, kp.to(cuda), None, 0.06, "linear", "sum",
                                       False, False, None, None, tr[0], tr[1])
        (out * gout.to(cuda)).sum().backward()
        grads.append((xg.grad.cpu(), Wg.grad.cpu()))
    assert rel_err(grads[1][0], x.grad) < TOL and rel_err(grads[1][1], W.grad) < TOL
    assert rel_err(grads[1][0], grads[0][0]) < 2e-5


@pytest.mark.parametrize("nq,ns,H,idx_dtype,strided", [(6000, 5000, 35, torch.int32, False), (2370, 9000, 35, torch.int64, True),
                                                       (16, 40, 48, torch.int32, False), (4737, 300, 8, torch.int64, False),
                                                       (1, 1, 1, torch.int32, False), (333, 500, 41, torch.int32, True)])
def test_fused_kernel_many_batches_vs_oracle(cuda, built_lib, nq, ns, H, idx_dtype, strided):
    """The fused kernel (gather + correlation + tcgen05 contraction, W^T in tensor memory) over several batches per CTA,
    partial last batches, shadow-only rows, int32/int64 and strided index rows: output with bias + LeakyReLU, 1/n and the
    optional wf against the CPU oracle; bit-identical run to run; wf == the two-kernel path's wf."""
    from oracle import model_ref
    from d3feat.pytorch_b200 import ops
    built_lib.d3f_set_kpconv_impl(3)
    try:
        assert built_lib.d3f_kpconv_fused_eligible(H, 15, 32, 32) == 1
        rng = np.random.default_rng(nq + H)
        q = torch.from_numpy((rng.random((nq, 3)) * 0.2).astype(np.float32))
        s = torch.from_numpy((rng.random((ns, 3)) * 0.2).astype(np.float32))
        wide = rng.integers(0, ns + 1, size=(nq, H + 5)).astype(np.int64)          # ns = shadow
        wide[rng.random(nq) < 0.05] = ns                                         # all-shadow rows
        wide = np.sort(wide, axis=1) if nq % 2 else wide                         # padding at the end (typical) or scattered
        inds = torch.from_numpy(wide)[:, :H] if strided else torch.from_numpy(np.ascontiguousarray(wide[:, :H]))
        x = torch.from_numpy(rng.standard_normal((ns, 32)).astype(np.float32))
        x[rng.choice(ns, max(ns // 10, 1), replace=False)] *= -1.0                 # rows with a non-positive channel sum
        W = torch.from_numpy((rng.standard_normal((15, 32, 32)) / np.sqrt(480)).astype(np.float32))
        kp = torch.from_numpy((_inputs.unit_kernel_points() * 0.075).astype(np.float32))
        b = torch.from_numpy(rng.standard_normal(32).astype(np.float32) * 0.1)
        ref = torch.nn.functional.leaky_relu(model_ref.kpconv_rigid(q, s, inds, x, W, kp, 0.06) + b, 0.1)
        g = lambda t: t.to(cuda)
        ig = g(torch.from_numpy(wide)).to(idx_dtype)
        ig = ig[:, :H] if strided else ig[:, :H].contiguous()
        args = (g(q), g(s), ig, g(x), g(W), g(kp), 0.06, "linear", "sum")
        out, wf, _, inv_n, _ = ops.kpconv_forward(*args, bias=g(b), slope=0.1, need_wf=True)
        out2, wf2, _, _, _ = ops.kpconv_forward(*args, bias=g(b), slope=0.1, need_wf=False)
        assert wf2 is None and torch.equal(out, out2)
        scale = max(float(ref.abs().max()), 1e-6)
        assert float((out.cpu() - ref).abs().max()) / scale < TOL
        built_lib.d3f_set_kpconv_impl(2)
        out3, wf3, _, inv3, _ = ops.kpconv_forward(*args, bias=g(b), slope=0.1)
        assert torch.equal(inv_n, inv3)
        assert float((wf - wf3).abs().max()) <= 1e-5 * max(float(wf3.abs().max()), 1e-6)
        assert built_lib.d3f_gemm_tcgen05_failed() == 0
    finally:
        built_lib.d3f_set_kpconv_impl(-1)


def test_transposed_lists_are_sorted_and_backward_is_reproducible(cuda):
    """d3f_neighbors_transpose lists queries in ascending order (also lists above the 256-entry shared-memory path), so
    the list-based grad_x is bit-identical run to run (atomic fill order)."""
    from d3feat.pytorch_b200 import ops
    rng = np.random.default_rng(5)
    for nq, ns, H in ((3000, 700, 35), (900, 2, 2)):          # in-degrees ~150 and 900 (> 256: the in-place path)
        inds = torch.from_numpy(np.stack([rng.permutation(ns + 1)[:H] for _ in range(nq)]).astype(np.int32)).to(cuda)
        t_off, t_src = ops.neighbors_transpose(inds, ns)
        off, src = t_off.cpu().numpy(), t_src.cpu().numpy()
        for j in range(ns):
            seg = src[off[j]:off[j + 1]]
            assert np.all(np.diff(seg) > 0), j
            assert np.array_equal(seg, np.nonzero((inds.cpu().numpy() == j).any(1))[0])
        again = ops.neighbors_transpose(inds, ns)
        assert torch.equal(again[0], t_off) and torch.equal(again[1][:off[-1]], t_src[:off[-1]])


@pytest.mark.parametrize("nq,ns,H,cin,cout,infl", [(300, 300, 60, 64, 64, "linear"), (90, 90, 151, 128, 32, "linear"),
                                                   (200, 150, 40, 32, 64, "gaussian")])
def test_deformable_backward_over_transposed_lists(cuda, nq, ns, H, cin, cout, infl):
    """Deformable layer (per-query kernel points + the in-range neighbour filter of blocks.py:300-324): grad_x through the
    gather over the transposed lists, grad_kernel_points from the scatter kernel without its grad_x reductions -- against
    the CPU oracle (autograd through oracle.model_ref.kpconv_rigid(kp_per_query=...)) and against the all-scatter path."""
    from oracle import model_ref
    from d3feat.pytorch_b200 import ops
    from d3feat.pytorch_b200.blocks import _KPConvFunction
    rng = np.random.default_rng(nq + H)
    q = torch.from_numpy((rng.random((nq, 3)) * 0.2).astype(np.float32))
    s = q.clone() if ns == nq else torch.from_numpy((rng.random((ns, 3)) * 0.2).astype(np.float32))
    inds = torch.from_numpy(rng.integers(0, ns + 1, size=(nq, H)).astype(np.int64))      # ns = shadow
    x = torch.from_numpy(rng.standard_normal((ns, cin)).astype(np.float32)).requires_grad_(True)
    W = torch.from_numpy((rng.standard_normal((15, cin, cout)) / np.sqrt(15 * cin)).astype(np.float32)).requires_grad_(True)
    kp0 = (_inputs.unit_kernel_points() * 0.075).astype(np.float32)
    kpd = torch.from_numpy(kp0[None] + 0.01 * rng.standard_normal((nq, 15, 3)).astype(np.float32)).requires_grad_(True)
    gout = torch.from_numpy(rng.standard_normal((nq, cout)).astype(np.float32))
    ref = model_ref.kpconv_rigid(q, s, inds, x, W, torch.from_numpy(kp0), 0.06, influence=infl, kp_per_query=kpd)
    (ref * gout).sum().backward()
    inds_g = inds.to(cuda).to(torch.int32)
    lists = ops.neighbors_transpose(inds_g, ns)
    grads = []
    for tr in ((None, None), lists):
        xg = x.detach().to(cuda).requires_grad_(True)
        Wg = W.detach().to(cuda).requires_grad_(True)
        kg = kpd.detach().to(cuda).requires_grad_(True)
        out, _ = _KPConvFunction.apply(q.to(cuda), s.to(cuda), inds_g, xg, Wg, kg, None, 0.06, infl, "sum",
                                       True, False, None, None, tr[0], tr[1])
        assert rel_err(out.detach().cpu(), ref.detach()) < TOL
        (out * gout.to(cuda)).sum().backward()
        grads.append((xg.grad.cpu(), Wg.grad.cpu(), kg.grad.cpu()))
    for got, want, name in zip(grads[1], (x.grad, W.grad, kpd.grad), ("x", "W", "kp")):
        assert rel_err(got, want) < TOL, name
    for a, b in zip(grads[1], grads[0]):
        assert rel_err(a, b) < 2e-5
