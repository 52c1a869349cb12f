"""Timeline of ONE replay of the CUDA-graph pair step from a torch.profiler trace: what runs when, how much of the step
some kernel is running at all (union of intervals), and where the idle gaps are.  python tools/step_timeline.py"""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import _inputs
from d3feat.pytorch_b200 import synthetic
from d3feat.pytorch_b200.architectures import KPFCNN
from d3feat.pytorch_b200.config import default_config
from d3feat.pytorch_b200.dataloader import collate_fn_descriptor
from d3feat.pytorch_b200.engine import PairStep, plan_capacities
from d3feat.pytorch_b200.loss import PairLoss
from d3feat.pytorch_b200.optim import FlatSGD

n = 20000
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
cfg = default_config()
model = KPFCNN(cfg).to(dev); model.load_state_dict(_inputs.kpfcnn_state_dict(cfg, seed=0)); model.train()
opt = FlatSGD(model)
pairs = [synthetic.fragment_pair(n, seed=i) for i in range(2)]
limits = [35, 42, 42, 45, 47]
sizes = [[int(t.shape[0]) for t in collate_fn_descriptor([p], cfg, limits)["points"]] for p in pairs]
st = PairStep(model, cfg, limits, plan_capacities(sizes), n, n, PairLoss("circle"), opt, None)
st(pairs[0]); st.capture()
dv = [tuple(torch.as_tensor(a).to(dev) for a in p) for p in pairs]
for i in range(3): st(dv[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    st(dv[0]); torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "step_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in ev)
print("kernels/memops: %d, span %.3f ms, sum of durations %.3f ms" % (len(ev), (t1 - t0) / 1e3, sum(e["dur"] for e in ev) / 1e3))
# union of busy intervals
busy, cur_s, cur_e = 0.0, None, None
for e in ev:
    s, f = e["ts"], e["ts"] + e["dur"]
    if cur_e is None or s > cur_e:
        if cur_e is not None: busy += cur_e - cur_s
        cur_s, cur_e = s, f
    else:
        cur_e = max(cur_e, f)
busy += cur_e - cur_s
print("some kernel running: %.3f ms (%.0f%% of the span); idle gaps: %.3f ms" % (busy / 1e3, 100 * busy / (t1 - t0), (t1 - t0 - busy) / 1e3))
def short(nm):
    nm = nm.replace("void ", "").replace("(anonymous namespace)::", "")
    return nm[:46]
bucket = 250.0
nb = int((t1 - t0) / bucket) + 1
for b in range(nb):
    lo, hi = t0 + b * bucket, t0 + (b + 1) * bucket
    acc = {}
    for e in ev:
        s, f = max(e["ts"], lo), min(e["ts"] + e["dur"], hi)
        if f > s: acc[short(e["name"])] = acc.get(short(e["name"]), 0.0) + (f - s)
    top = sorted(acc.items(), key=lambda t: -t[1])[:4]
    print("%5.2f ms | occupancy %.2f |" % (b * bucket / 1e3, sum(acc.values()) / bucket), "; ".join("%s %.0f" % (k, v) for k, v in top))
out = os.path.join(ROOT, "gpurun_out", "step_events.tsv")
os.makedirs(os.path.dirname(out), exist_ok=True)
with open(out, "w") as f:
    for e in ev:
        a = e.get("args", {})
        f.write("%.1f\t%.1f\t%s\t%s\t%s\n" % (e["ts"] - t0, e["dur"], a.get("stream", "?"), a.get("grid", ""), short(e["name"])))
print("wrote", out)
