"""Sweep tile width / K splits of the tcgen05 GEMM over every GEMM shape of the pair step:
       python tools/gemm_tune.py [n_points]
Logs the (M, N, K, transA, transB, mode) of every ops.gemm call of one eager training step, then times each unique plain
shape warm (operands L2-resident, as inside the step where the producer has just written them; a spin kernel queued in
front of each call hides the host's launch latency from the events) under the heuristic and
under forced (BN, splits).  Output: one line per shape with the heuristic time, the best forced setting and its time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from d3feat.pytorch_b200 import synthetic, ops, _lib
from d3feat.pytorch_b200.architectures import KPFCNN
from d3feat.pytorch_b200.config import default_config
from d3feat.pytorch_b200.dataloader import calibrate_neighbors, collate_fn_descriptor
from d3feat.pytorch_b200.engine import PairStep, plan_capacities
from d3feat.pytorch_b200.loss import PairLoss
from d3feat.pytorch_b200.optim import FlatSGD

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
torch.cuda.set_device(0); dev = torch.device("cuda:0")
cfg = default_config(); torch.manual_seed(0); np.random.seed(0)
model = KPFCNN(cfg).to(dev); model.train()
opt = FlatSGD(model, lr=0.01, momentum=0.98, weight_decay=1e-6)
pairs = [synthetic.fragment_pair(n, seed=i) for i in range(2)]
class DS:
    config = cfg
    def __len__(self): return 2
    def __getitem__(self, i): return pairs[i]
limits = [int(v) for v in calibrate_neighbors(DS(), cfg, collate_fn_descriptor, samples_threshold=10 ** 9)]
sizes = [[int(t.shape[0]) for t in collate_fn_descriptor([p], cfg, limits)["points"]] for p in pairs]
st = PairStep(model, cfg, limits, plan_capacities(sizes), n, n, PairLoss("circle"), opt, None)

log = {}
real_gemm = ops.gemm
def logged(a, b, trans_a=False, trans_b=False, row_scale=None, k_scale=None, bias=None, slope=None, deterministic=False,
           bias2=None, residual=None, out=None):
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    N = b.shape[0] if trans_b else b.shape[1]
    mode = "det" if (deterministic or bias2 is not None or residual is not None) else "plain"
    key = (M, N, K, bool(trans_a), bool(trans_b), mode, bias is not None or slope is not None)
    log[key] = log.get(key, 0) + 1
    return real_gemm(a, b, trans_a, trans_b, row_scale, k_scale, bias, slope, deterministic, bias2, residual, out)
ops.gemm = logged
st(pairs[0]); torch.cuda.synchronize()
ops.gemm = real_gemm
lib = _lib.load()

def time_one(M, N, K, ta, tb, mode, epi, reps=30):
    a = torch.randn((K, M) if ta else (M, K), device=dev)
    b = torch.randn((N, K) if tb else (K, N), device=dev)
    bias = torch.randn(N, device=dev) if epi else None
    kw = dict(trans_a=ta, trans_b=tb, bias=bias, slope=0.1 if epi else None, deterministic=(mode == "det"))
    for _ in range(3): ops.gemm(a, b, **kw)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(300000)      # keep the GPU busy while the host enqueues: the events then bracket GPU time only
        e0.record(); ops.gemm(a, b, **kw); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))

print("%-44s %5s %9s   %s" % ("M N K tA tB mode epi", "calls", "heur us", "forced (bn, splits): us ..."))
tot_h = tot_b = 0.0
for key, calls in sorted(log.items(), key=lambda kv: -kv[0][0] * kv[0][1] * kv[0][2]):
    M, N, K, ta, tb, mode, epi = key
    lib.d3f_set_gemm_tuning(0, 0)
    h = time_one(*key)
    res = []
    for bn in ((32,) if N <= 32 else (64, 128)):
        for sp in ((1, 2, 3, 4, 6, 8, 12, 16, 24) if (mode == "plain" and not epi) else (0,)):
            if sp > 1 and K < sp * 64: continue
            lib.d3f_set_gemm_tuning(bn, sp)
            res.append((time_one(*key, reps=15), bn, sp))
    lib.d3f_set_gemm_tuning(0, 0)
    res.sort()
    best = res[0]
    tot_h += h * calls; tot_b += min(h, best[0]) * calls
    print("%-44s %5d %9.1f   best %s %.1f | %s" % (" ".join(map(str, key)), calls, h, best[1:], best[0],
          "  ".join("%d/%d:%.0f" % (bn, sp, t) for t, bn, sp in sorted(res, key=lambda r: (r[1], r[2])))))
print("sum over calls: heuristic %.0f us, best-forced %.0f us" % (tot_h, tot_b))
