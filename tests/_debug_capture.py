import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import _inputs
from d3feat.pytorch_b200 import synthetic, parallel
from d3feat.pytorch_b200.architectures import KPFCNN
from d3feat.pytorch_b200.config import default_config
from d3feat.pytorch_b200.dataloader import collate_fn_descriptor
from d3feat.pytorch_b200.engine import PairStep, plan_capacities, collate_static
from d3feat.pytorch_b200.loss import PairLoss
torch.cuda.set_device(0)
dev = torch.device("cuda:0")
cfg = default_config(first_features_dim=32, num_node=64)
limits = [40]*5
model = KPFCNN(cfg).to(dev); model.train()
loss_fn = PairLoss("circle")
data = synthetic.fragment_pair(1500, seed=5, num_node=64)
sizes = [[int(p.shape[0]) for p in collate_fn_descriptor([data], cfg, limits)["points"]]]
caps = plan_capacities(sizes, margin=1.2, align=32)
print("caps", caps, flush=True)

def try_capture(name, fn):
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2): fn()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("CAPTURE OK:", name, flush=True)
    except Exception:
        print("CAPTURE FAILED:", name, flush=True)
        traceback.print_exc()
        try: torch.cuda.synchronize()
        except Exception: pass

dev_in = [torch.as_tensor(a).to(dev) for a in data]
len0 = torch.tensor([1500,1500], dtype=torch.int32, device=dev)
from d3feat.pytorch_b200 import ops
try_capture("radius only", lambda: ops.radius_neighbors_raw(dev_in[0], dev_in[0], len0[:1], len0[:1], 0.075, 40, torch.int32, None, False, pad_index=1500))
try_capture("subsample only", lambda: ops.grid_subsample_raw(dev_in[0], len0[:1], 0.06, 1500))
try_capture("collate", lambda: collate_static(*dev_in, cfg, limits, caps, len0))
batch, st = collate_static(*dev_in, cfg, limits, caps, len0)
def fwd():
    with torch.no_grad():
        return model(batch)
try_capture("forward", fwd)
opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.98, weight_decay=1e-6)
flat = parallel.FlatGradients(model)
stp = PairStep(model, cfg, limits, caps, 1500, 1500, loss_fn, opt, flat, num_node=64)
stp.load(data)
try_capture("full step", stp._body)
