import sys, os, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import _inputs
from d3feat.pytorch_b200 import synthetic
from d3feat.pytorch_b200.architectures import KPFCNN
from d3feat.pytorch_b200.blocks import gather
from d3feat.pytorch_b200.config import default_config
from d3feat.pytorch_b200.dataloader import collate_fn_descriptor
from d3feat.pytorch_b200.engine import PairStep, plan_capacities
from d3feat.pytorch_b200.loss import PairLoss
from oracle import model_ref, pipeline
dev = torch.device("cuda:0")
cfg = default_config(first_features_dim=32, num_node=64); limits=[40]*5
sd = _inputs.kpfcnn_state_dict(cfg, seed=3)
model = KPFCNN(cfg).to(dev); model.load_state_dict(sd); model.train()
loss_fn = PairLoss("circle", "euclidean", 10, 0.1, 0.1, 1.4)
data = synthetic.fragment_pair(1500, seed=5, num_node=64)
def dropin(idt=torch.int64):
    batch = collate_fn_descriptor([data], cfg, limits, index_dtype=idt)
    feats, scores = model(batch)
    c = batch["corr"].long(); ia, ip = c[:,0], c[:,1]+1500
    o = loss_fn(gather(feats, ia), gather(feats, ip), batch["dist_keypts"], gather(scores, ia), gather(scores, ip))
    model.zero_grad(set_to_none=True)
    (o["desc_loss"]+o["det_loss"]).backward()
    return {k: p.grad.clone().cpu() for k,p in model.named_parameters() if p.grad is not None}, feats.detach().cpu()
def cmp(a, b, name):
    worst = sorted(((float((a[k]-b[k]).abs().max()/max(float(b[k].abs().max()),1e-30)), k) for k in a), reverse=True)[:4]
    print(name, ["%.1e %s" % w for w in worst], flush=True)
g1, f1 = dropin(); g2, f2 = dropin(); g3, f3 = dropin(torch.int32)
print("fwd run-to-run", float((f1-f2).abs().max()), "int32 vs int64", float((f1-f3).abs().max()))
cmp(g1, g2, "dropin run-to-run")
cmp(g3, g1, "dropin int32 vs int64")
sizes = [[int(p.shape[0]) for p in collate_fn_descriptor([data], cfg, limits)["points"]]]
for margin in (1.0, 1.2):
    caps = plan_capacities(sizes, margin=margin, align=1 if margin==1.0 else 32)
    class NoOpt:
        def zero_grad(self, set_to_none=False):
            for p in model.parameters():
                if p.grad is not None: p.grad.zero_()
        def step(self): pass
    for p in model.parameters(): p.grad = torch.zeros_like(p)
    st = PairStep(model, cfg, limits, caps, 1500, 1500, loss_fn, NoOpt(), None, num_node=64)
    st(data); st.check()
    gs = {k: p.grad.clone().cpu() for k,p in model.named_parameters()}
    cmp({k: gs[k] for k in g1}, g1, "static caps=%s vs dropin" % caps)
# oracle
cb = pipeline.cpu_collate(data, cfg, limits, impl="port")
params = {k: v.clone().requires_grad_("kernel_points" not in k) for k,v in sd.items()}
f, s = model_ref.kpfcnn_forward(params, cb, cfg, training=True)
dl, det, _, _ = model_ref.pair_losses(f, s, cb, "circle"); (dl+det).backward()
go = {k: params[k].grad for k in g1}
cmp(g1, go, "dropin vs CPU oracle")
