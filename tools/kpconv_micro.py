"""A/B micro-benchmark of the KPConv gather / scatter kernels and the contraction GEMM on the real pyramid of one
synthetic 20k+20k pair: python tools/kpconv_micro.py   (CUDA events around `reps` back-to-back calls, L2-warm)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from d3feat.pytorch_b200 import _lib, ops, synthetic
from d3feat.pytorch_b200.config import default_config
from d3feat.pytorch_b200.dataloader import collate_fn_descriptor

lib = _lib.load()
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
cfg = default_config()
limits = [35, 42, 42, 45, 47]
batch = collate_fn_descriptor([synthetic.fragment_pair(20000, seed=0)], cfg, limits)
kp = torch.from_numpy(np.random.default_rng(0).standard_normal((15, 3)).astype(np.float32) * 0.03).to(dev)


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


cases = [("L0 32->32", 0, "neighbors", 32, 32), ("L0->1 strided 32->32", 0, "pools", 32, 32), ("L1 64->64", 1, "neighbors", 64, 64),
         ("L2 128->128", 2, "neighbors", 128, 128), ("L0 1->64", 0, "neighbors", 1, 64)]
print("%-22s %-10s %10s %10s %10s" % ("layer", "impl", "gather us", "fwd op us", "bwd op us"))
for name, lvl, kind, cin, cout in cases:
    inds = batch[kind][lvl].to(torch.int32)
    s = batch["points"][lvl]
    q = batch["points"][lvl + 1] if kind == "pools" else s
    r = cfg.first_subsampling_dl * cfg.conv_radius * (2 ** lvl)
    x = torch.randn(s.shape[0], cin, device=dev)
    W = torch.randn(15, cin, cout, device=dev) / (15 * cin) ** 0.5
    kpl = kp * (2 ** lvl)
    g = torch.randn(q.shape[0], cout, device=dev)
    tr = ops.neighbors_transpose(inds, s.shape[0])
    t_tr = timed(lambda: ops.neighbors_transpose(inds, s.shape[0]))
    print("%-22s neighbors_transpose: %.1f us" % (name, t_tr))
    for impl, label, env in ((0, "v1", None), (1, "v2-ffma", None), (2, "v2-mma", None), (2, "v2-mma/scalar-red", "scalar"),
                             (2, "v2-mma/transposed", "t"), (2, "v2-mma/transp+skinny", "sk")):
        lib.d3f_set_gemm_skinny(2 if env == "sk" else 0)
        lib.d3f_set_kpconv_impl(impl)
        if hasattr(lib, "d3f_set_scatter_vec"):
            lib.d3f_set_scatter_vec(0 if env == "scalar" else 2)
        elif env == "scalar":
            continue
        ext = 0.8 * r
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record(); ev[1].record()
        out = ops.kpconv_forward(q, s, inds, x, W, kpl, ext, "linear", "sum")
        t_f = timed(lambda: ops.kpconv_forward(q, s, inds, x, W, kpl, ext, "linear", "sum"))
        # gather kernel alone: events recorded inside the C call (last call of a short loop)
        lib.d3f_kpconv_set_gather_events(ev[0].cuda_event, ev[1].cuda_event)
        for _ in range(3):
            ops.kpconv_forward(q, s, inds, x, W, kpl, ext, "linear", "sum")
        torch.cuda.synchronize()
        lib.d3f_kpconv_set_gather_events(None, None)
        t_g = ev[0].elapsed_time(ev[1]) * 1e3
        _, wf, wf_un, inv_n, _ = out
        t_b = timed(lambda: ops.kpconv_backward(q, s, inds, x, W, kpl, ext, "linear", "sum", False, None, wf, wf_un, inv_n, g,
                                                cin > 1, True, False, False, transpose=tr if env in ("t", "sk") else None))
        print("%-22s %-18s %10.1f %10.1f %10.1f" % (name, label, t_g, t_f, t_b))
lib.d3f_set_kpconv_impl(-1)
lib.d3f_set_gemm_skinny(-1)

print("\nGEMM pipelines (us per call, back to back):")
shapes = [(40000, 32, 480, False, False), (13312, 32, 480, False, False), (13312, 64, 960, False, False), (40000, 480, 32, False, True),
          (480, 32, 40000, True, False), (40000, 128, 32, False, True), (40000, 32, 128, False, True), (40000, 32, 384, False, True),
          (2816, 128, 1920, False, False), (768, 1024, 3072, False, True), (256, 512, 7680, False, False), (7680, 512, 256, True, False)]
for (M, N, K, ta, tb) in shapes:
    a = torch.randn((K, M) if ta else (M, K), device=dev)
    b = torch.randn((N, K) if tb else (K, N), device=dev)
    row = []
    pipes = (0, 1, 2, 3) if os.environ.get("D3F_MICRO_TMEM", "1") == "1" else (0, 1, 2)   # 3 = experimental A-in-TMEM kernel (tc8)
    for pipe in (0, 1, 2):
        lib.d3f_set_gemm_pipeline(pipe)
        for det in (False, True):
            row.append(timed(lambda: ops.gemm(a, b, ta, tb, deterministic=det)))
    lib.d3f_set_gemm_pipeline(-1)
    lib.d3f_set_gemm_skinny(2)
    row.append(timed(lambda: ops.gemm(a, b, ta, tb, deterministic=True)))
    lib.d3f_set_gemm_skinny(-1)
    t8 = None
    if 3 in pipes:
        lib.d3f_set_gemm_pipeline(3)
        t8 = (timed(lambda: ops.gemm(a, b, ta, tb, deterministic=False)), timed(lambda: ops.gemm(a, b, ta, tb, deterministic=True)))
        lib.d3f_set_gemm_pipeline(-1)
    fl = 2.0 * M * N * K
    print("M=%6d N=%5d K=%6d ta=%d tb=%d | reg: %6.1f (det %6.1f) | cp.async: %6.1f (det %6.1f) | warp-spec: %6.1f (det %6.1f) | skinny(if eligible): %6.1f | %.1f TFLOP/s best"
          % (M, N, K, ta, tb, row[0], row[1], row[2], row[3], row[4], row[5], row[6], fl / min(row) / 1e6)
          + ("" if t8 is None else " | tmem-A (tc8): %6.1f (det %6.1f)" % t8))
if lib.d3f_gemm_tcgen05_failed() != 0:
    print("WARNING: a tcgen05 GEMM gave up waiting on an mbarrier")
