"""The 3xTF32 tensor-core GEMM must be fp32-accurate (plain TF32 would eat the 1e-4 budget)."""
import numpy as np
import pytest
import torch

from _util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tcgen05", "mma"])
def gemm_impl(request, cuda):
    """tcgen05 = the product kernel (tcgen05.mma, accumulators in TMEM); mma = legacy mma.sync kernel (debug selector,
    include/d3feat_b200_debug.h).  Round 2 deleted the cp.async / warp-specialised / TMEM-A variants after timing them."""
    from d3feat.pytorch_b200 import _lib
    lib = _lib.load()
    lib.d3f_set_gemm_impl(0 if request.param == "mma" else 1)
    yield request.param
    if request.param != "mma":
        assert lib.d3f_gemm_tcgen05_failed() == 0, "a tcgen05 GEMM gave up waiting on its mbarrier"
    lib.d3f_set_gemm_impl(1)


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (127, 65, 33), (128, 64, 32), (300, 45, 15), (1000, 512, 960),
                                   (190, 512, 7680), (480, 32, 40000), (4097, 33, 130), (4100, 32, 480), (132, 128, 36),
                                   (260, 96, 100), (13312, 64, 960)])
def test_gemm_matches_fp64(cuda, gemm_impl, ta, tb, M, N, K):
    from d3feat.pytorch_b200 import ops
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    rs = (rng.random(M) + 0.5).astype(np.float32)
    ks = (rng.random(K) + 0.5).astype(np.float32) if not tb else None
    opA = A.T.astype(np.float64) if ta else A.astype(np.float64)
    opB = B.T.astype(np.float64) if tb else B.astype(np.float64)
    ref = rs[:, None] * (opA @ ((ks[:, None].astype(np.float64) if ks is not None else 1.0) * opB))
    got = ops.gemm(torch.from_numpy(A).to(cuda), torch.from_numpy(B).to(cuda), ta, tb,
                   row_scale=torch.from_numpy(rs).to(cuda), k_scale=None if ks is None else torch.from_numpy(ks).to(cuda))
    # fp32 accuracy: error relative to |A||B| row/col magnitudes ~ sqrt(K) * 2^-23
    assert rel_err(got.cpu(), ref) < 2e-6 * max(1.0, np.sqrt(K) / 8), (M, N, K)


def test_gemm_handles_strided_rows_and_bias_activation(cuda, gemm_impl):
    from d3feat.pytorch_b200 import ops
    rng = np.random.default_rng(3)
    X = torch.from_numpy(rng.standard_normal((500, 70)).astype(np.float32)).to(cuda)
    W = torch.from_numpy(rng.standard_normal((48, 70)).astype(np.float32)).to(cuda)
    b = torch.from_numpy(rng.standard_normal(48).astype(np.float32)).to(cuda)
    ref = torch.nn.functional.leaky_relu(X.double() @ W.double().t() + b.double(), 0.1)
    got = ops.gemm(X, W, trans_b=True, bias=b, slope=0.1)
    assert rel_err(got.cpu(), ref.cpu()) < 2e-6


def test_fused_linear_autograd(cuda, gemm_impl):
    from d3feat.pytorch_b200 import ops
    rng = np.random.default_rng(4)
    x = torch.from_numpy(rng.standard_normal((777, 96)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((40, 96)) / 10).astype(np.float32))
    b = torch.from_numpy(rng.standard_normal(40).astype(np.float32))
    g = torch.from_numpy(rng.standard_normal((777, 40)).astype(np.float32))
    for slope in (0.1, None):
        xr, wr, br = (t.double().requires_grad_(True) for t in (x, w, b))
        y = xr @ wr.t() + br
        if slope is not None:
            y = torch.nn.functional.leaky_relu(y, slope)
        (y * g.double()).sum().backward()
        xg, wg, bg = (t.to(cuda).requires_grad_(True) for t in (x, w, b))
        out = ops.fused_linear(xg, wg, bg, slope)
        (out * g.to(cuda)).sum().backward()
        assert rel_err(out.detach().cpu(), y.detach()) < 2e-6
        assert rel_err(xg.grad.cpu(), xr.grad) < 2e-6 and rel_err(wg.grad.cpu(), wr.grad) < 5e-6
        assert rel_err(bg.grad.cpu(), br.grad) < 2e-6


@pytest.mark.parametrize("M,N,K", [(256, 512, 2048), (693, 1024, 3072), (13312, 64, 960), (100, 40, 511), (100, 40, 512)])
def test_gemm_deterministic_split_with_epilogue(cuda, gemm_impl, M, N, K):
    """d3f_gemm_ex: K-only split rule + ordered partial sums; bias / LeakyReLU applied after the reduction, and a
    row's bits do not depend on how many rows the problem has (static-capacity padding)."""
    from d3feat.pytorch_b200 import ops
    rng = np.random.default_rng(M + N + K)
    X = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float32)).to(cuda)
    W = torch.from_numpy((rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)).to(cuda)
    b = torch.from_numpy(rng.standard_normal(N).astype(np.float32)).to(cuda)
    ref = torch.nn.functional.leaky_relu(X.double() @ W.double().t() + b.double(), 0.1)
    got = ops.gemm(X, W, trans_b=True, bias=b, slope=0.1, deterministic=True)
    assert rel_err(got.cpu(), ref.cpu()) < 2e-6 * max(1.0, np.sqrt(K) / 8)
    again = ops.gemm(X, W, trans_b=True, bias=b, slope=0.1, deterministic=True)
    assert torch.equal(got, again)
    part = ops.gemm(X[: M // 2 + 1].contiguous(), W, trans_b=True, bias=b, slope=0.1, deterministic=True)
    assert torch.equal(got[: M // 2 + 1], part)


def test_fused_linear_two_biases_residual_autograd(cuda, gemm_impl):
    """leaky(x W^T + b + b2 + residual): the ResnetBottleneckBlock tail as one GEMM (blocks.py:686)."""
    from d3feat.pytorch_b200 import ops
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal((333, 64)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((256, 64)) / 8).astype(np.float32))
    b = torch.from_numpy(rng.standard_normal(256).astype(np.float32))
    b2 = torch.from_numpy(rng.standard_normal(256).astype(np.float32))
    r = torch.from_numpy(rng.standard_normal((333, 256)).astype(np.float32))
    g = torch.from_numpy(rng.standard_normal((333, 256)).astype(np.float32))
    ref_in = [t.double().requires_grad_(True) for t in (x, w, b, b2, r)]
    y = torch.nn.functional.leaky_relu(ref_in[0] @ ref_in[1].t() + ref_in[2] + ref_in[3] + ref_in[4], 0.1)
    (y * g.double()).sum().backward()
    gpu_in = [t.to(cuda).requires_grad_(True) for t in (x, w, b, b2, r)]
    out = ops.fused_linear(gpu_in[0], gpu_in[1], gpu_in[2], 0.1, bias2=gpu_in[3], residual=gpu_in[4])
    (out * g.to(cuda)).sum().backward()
    assert rel_err(out.detach().cpu(), y.detach()) < 2e-6
    for a, c in zip(gpu_in, ref_in):
        assert rel_err(a.grad.cpu(), c.grad) < 5e-6


@pytest.mark.parametrize("M,N", [(40000, 32), (13312, 64), (777, 128), (5, 2048), (1, 4), (1000, 40), (333, 7), (4097, 512)])
def test_colsum(cuda, M, N):
    from d3feat.pytorch_b200 import ops
    rng = np.random.default_rng(M + N)
    x = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32)).to(cuda)
    ref = x.double().sum(0)
    got = ops.colsum(x)
    assert float((got.double() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("M,N", [(40000, 32), (2816, 128), (5, 2048), (1000, 40), (333, 64)])
def test_leaky_backward_colsum(cuda, M, N):
    from d3feat.pytorch_b200 import ops
    rng = np.random.default_rng(M * 3 + N)
    g = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32)).to(cuda)
    y = torch.from_numpy(rng.standard_normal((M, N)).astype(np.float32)).to(cuda)
    dz, db = ops.leaky_backward_colsum(g, y, 0.1)
    ref = g * torch.where(y > 0, 1.0, 0.1)
    assert torch.equal(dz, ref)
    assert float((db.double() - ref.double().sum(0)).abs().max()) < 2e-5 * max(1.0, float(ref.double().sum(0).abs().max()))


def test_gemm_blocked_weight_transpose(cuda):
    """The KPConv data-gradient GEMM: dx = G [Ns, K*Cout] x W^T with W [K, Cin, Cout] addressed block-wise."""
    rng = np.random.default_rng(9)
    ns, Kp, cin, cout = 5000, 15, 32, 64
    G = torch.from_numpy(rng.standard_normal((ns, Kp * cout)).astype(np.float32)).to(cuda)
    W = torch.from_numpy((rng.standard_normal((Kp, cin, cout)) / 30).astype(np.float32)).to(cuda)
    ref = G.double() @ W.double().permute(0, 2, 1).reshape(Kp * cout, cin)
    from d3feat.pytorch_b200 import _lib
    lib = _lib.load()
    out = torch.empty((ns, cin), dtype=torch.float32, device=cuda)
    # through the public path: a KPConv backward over transposed lists is covered in test_gpu_kpconv; here only the
    # GEMM addressing is checked against an explicitly transposed copy of the weights
    from d3feat.pytorch_b200 import ops
    Wt = W.permute(0, 2, 1).reshape(Kp * cout, cin).contiguous()
    got = ops.gemm(G, Wt)
    assert rel_err(got.cpu(), ref.cpu()) < 2e-6 * np.sqrt(Kp * cout) / 8
