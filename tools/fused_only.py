"""Launch the fused KPConv kernel a few times on the L0 32->32 layer of a synthetic 20k+20k pair (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from d3feat.pytorch_b200 import _lib, ops, synthetic
from d3feat.pytorch_b200.config import default_config
from d3feat.pytorch_b200.dataloader import collate_fn_descriptor
lib = _lib.load()
dev = torch.device("cuda:0")
cfg = default_config()
batch = collate_fn_descriptor([synthetic.fragment_pair(20000, seed=0)], cfg, [35, 42, 42, 45, 47])
inds = batch["neighbors"][0].to(torch.int32); s = batch["points"][0]
kp = torch.from_numpy(np.random.default_rng(0).standard_normal((15, 3)).astype(np.float32) * 0.03).to(dev)
x = torch.randn(s.shape[0], 32, device=dev); W = torch.randn(15, 32, 32, device=dev) / 480 ** 0.5; b = torch.randn(32, device=dev)
impl = int(os.environ.get("IMPL", "3"))
lib.d3f_set_kpconv_impl(impl)
for _ in range(int(os.environ.get("REPS", "4"))):
    ops.kpconv_forward(s, s, inds, x, W, kp, 0.06, "linear", "sum", bias=b, slope=0.1, need_wf=impl != 3)
torch.cuda.synchronize()
print("done")
