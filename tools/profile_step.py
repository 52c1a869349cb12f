"""Kernel-time table of the CUDA-graph pair step (torch.profiler / CUPTI): python tools/profile_step.py [n_points] [deform]
(`deform`: BASELINE config 4 architecture -- deformable KPConv from level 3)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from d3feat.pytorch_b200 import synthetic, parallel
from d3feat.pytorch_b200.architectures import KPFCNN
from d3feat.pytorch_b200.config import default_config
from d3feat.pytorch_b200.dataloader import calibrate_neighbors, collate_fn_descriptor
from d3feat.pytorch_b200.engine import PairStep, plan_capacities
from d3feat.pytorch_b200.loss import PairLoss

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
torch.cuda.set_device(0); dev = torch.device("cuda:0")
cfg = default_config()
if len(sys.argv) > 2 and sys.argv[2] == "deform":
    from d3feat.pytorch_b200.config import build_architecture
    cfg = default_config(architecture=build_architecture(5, deformable_from=3))
torch.manual_seed(0); np.random.seed(0)
model = KPFCNN(cfg).to(dev); model.train()
from d3feat.pytorch_b200.optim import FlatSGD
opt = FlatSGD(model, lr=0.01, momentum=0.98, weight_decay=1e-6)
pairs = [synthetic.fragment_pair(n, seed=i) for i in range(2)]
class DS:
    config = cfg
    def __len__(self): return 2
    def __getitem__(self, i): return pairs[i]
limits = [int(v) for v in calibrate_neighbors(DS(), cfg, collate_fn_descriptor, samples_threshold=10 ** 9)]
sizes = [[int(t.shape[0]) for t in collate_fn_descriptor([p], cfg, limits)["points"]] for p in pairs]
st = PairStep(model, cfg, limits, plan_capacities(sizes), n, n, PairLoss("circle"), opt, None)
st(pairs[0]); st.capture()
for i in range(3): st(pairs[i % 2])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dv = [tuple(torch.as_tensor(a).to(dev) for a in p) for p in pairs]
e0.record()
for i in range(20): st(dv[i % 2])
e1.record(); torch.cuda.synchronize()
print("graph step (device-resident inputs, no L2 flush): %.3f ms" % (e0.elapsed_time(e1) / 20))
if os.environ.get("D3F_NCU"):   # under `ncu --profile-from-start off`: exactly one replay of the graph step is profiled
    torch.cuda.profiler.start()
    st(pairs[0]); torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(5): st(pairs[i % 2])
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 5.0, e.count / 5.0) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("total kernel time per step: %.3f ms over %.0f launches" % (tot / 1e3, sum(r[2] for r in rows)))
for k, t, c in rows[:45]:
    print("%8.1f us %5.1f%% n=%5.1f  %s" % (t, 100 * t / tot, c, k[:110]))
