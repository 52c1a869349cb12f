"""A/B micro-benchmark of the KPConv forward paths on the real pyramid of one synthetic 20k+20k pair:
python tools/kpconv_micro.py   -> per layer: gather/fused kernel alone (events inside the C call) and the whole op,
L2-warm (back to back) and L2-cold (256 MiB flush write before every call)."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20, cold=False):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    evs = []
    for _ in range(reps):
        if cold:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / reps * 1e3   # us


def kernel_alone(fn, reps=10, cold=False):
    ts = []
    for _ in range(reps):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record(); ev[1].record()
        if cold:
            flush.fill_(1)
        lib.d3f_kpconv_set_gather_events(ev[0].cuda_event, ev[1].cuda_event)
        fn()
        lib.d3f_kpconv_set_gather_events(None, None)
        torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]) * 1e3)
    return float(np.median(ts))


cases = [("L0 32->32", 0, "neighbors", 32, 32), ("L0->1 strided 32->32", 0, "pools", 32, 32), ("L1 64->64", 1, "neighbors", 64, 64),
         ("L2 128->128", 2, "neighbors", 128, 128), ("L0 1->64", 0, "neighbors", 1, 64)]
print("%-22s %-12s %12s %12s %12s %12s  %s" % ("layer", "impl", "kern warm", "kern cold", "op warm", "op cold", "logical GB/s (kernel, cold)"))
for name, lvl, kind, cin, cout in cases:
    inds = batch[kind][lvl].to(torch.int32)
    s = batch["points"][lvl]
    q = batch["points"][lvl + 1] if kind == "pools" else s
    r = cfg.first_subsampling_dl * cfg.conv_radius * (2 ** lvl)
    x = torch.randn(s.shape[0], cin, device=dev)
    W = torch.randn(15, cin, cout, device=dev) / (15 * cin) ** 0.5
    b = torch.randn(cout, device=dev)
    kpl = kp * (2 ** lvl)
    ext = 0.8 * r
    nq, H = q.shape[0], inds.shape[1]
    logical = nq * H * (4 * cin + 16) + nq * (12 + 4 * cout) + 4 * 15 * cin * cout
    outs = {}
    for impl, label in ((1, "ffma+gemm"), (2, "mma+gemm"), (3, "fused")):
        lib.d3f_set_kpconv_impl(impl)
        fused = impl == 3 and lib.d3f_kpconv_fused_eligible(H, 15, cin, cout)
        if impl == 3 and not fused:
            continue
        fn = lambda: ops.kpconv_forward(q, s, inds, x, W, kpl, ext, "linear", "sum", bias=b, slope=0.1, need_wf=not fused)
        outs[impl] = fn()[0]
        kw, kc = kernel_alone(fn), kernel_alone(fn, cold=True)
        ow, oc = timed(fn), timed(fn, cold=True)
        print("%-22s %-12s %12.1f %12.1f %12.1f %12.1f  %8.0f" % (name, label, kw, kc, ow, oc, logical / kc / 1e3))
    if 3 in outs:
        d = (outs[3] - outs[2]).abs().max() / outs[2].abs().max()
        print("%-22s fused vs mma+gemm max rel diff %.2e" % (name, float(d)))
lib.d3f_set_kpconv_impl(-1)
if lib.d3f_gemm_tcgen05_failed() != 0:
    print("WARNING: a tensor-core kernel gave up waiting on an mbarrier")
