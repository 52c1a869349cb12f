"""Launch d3f_grid_subsample on the level-0 cloud of a synthetic 20k+20k pair a few times (for ncu / timing)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from d3feat.pytorch_b200 import ops, synthetic
d = synthetic.fragment_pair(20000, seed=0)
pts = torch.from_numpy(np.concatenate([d[0], d[1]])).cuda()
lens = torch.tensor([20000, 20000], dtype=torch.int32, device="cuda")
for _ in range(3):
    out, ol = ops.grid_subsample_raw(pts, lens, 0.06, 14000)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out, ol = ops.grid_subsample_raw(pts, lens, 0.06, 14000)
e1.record(); torch.cuda.synchronize()
print("grid_subsample 40000 -> %s: %.1f us per call" % (ol.tolist(), e0.elapsed_time(e1) * 100))
