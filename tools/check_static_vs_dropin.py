"""Diagnostic: forward bit-equality and gradient agreement of engine.PairStep (static shapes) vs the drop-in pipeline."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import _inputs
from d3feat.pytorch_b200 import synthetic
from d3feat.pytorch_b200.architectures import KPFCNN
from d3feat.pytorch_b200.blocks import gather
from d3feat.pytorch_b200.config import default_config
from d3feat.pytorch_b200.dataloader import collate_fn_descriptor
from d3feat.pytorch_b200.engine import PairStep, plan_capacities, collate_static
from d3feat.pytorch_b200.loss import PairLoss
dev = torch.device("cuda:0")
cfg = default_config(first_features_dim=32, num_node=64); limits = [40] * 5
model = KPFCNN(cfg).to(dev); model.load_state_dict(_inputs.kpfcnn_state_dict(cfg, seed=3)); model.train()
loss_fn = PairLoss("circle", "euclidean", 10, 0.1, 0.1, 1.4)
data = synthetic.fragment_pair(1500, seed=5, num_node=64)
acts = {}
def hook(name):
    def f(mod, inp, out):
        if isinstance(out, torch.Tensor): acts.setdefault(name, []).append(out.detach().clone())
    return f
hs = [m.register_forward_hook(hook(n)) for n, m in model.named_modules() if n and n.count(".") <= 3]
def run(batch):
    feats, scores = model(batch)
    c = batch["corr"].long(); ia, ip = c[:, 0], c[:, 1] + 1500
    o = loss_fn(gather(feats, ia), gather(feats, ip), batch["dist_keypts"], gather(scores, ia), gather(scores, ip))
    model.zero_grad(set_to_none=True)
    (o["desc_loss"] + o["det_loss"]).backward()
    return {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
exact = collate_fn_descriptor([data], cfg, limits)
sizes = [int(p.shape[0]) for p in exact["points"]]
g_a = run(exact); g_a2 = run(exact)
caps = plan_capacities([sizes], margin=1.2, align=32)
static, _pyr = collate_static(*[torch.as_tensor(a).to(dev) for a in data], cfg, limits, caps)
status = _pyr.join()
g_b = run(static); g_b2 = run(static)
for h in hs: h.remove()
lvl_of = {}
print("sizes", sizes, "caps", caps)
bad = 0
for name, lst in acts.items():
    a, b = lst[0], lst[2]
    n = min(a.shape[0], b.shape[0])
    # real rows of a static tensor = first `size` rows at that level
    real = next((s for s, c in zip(sizes, caps) if c == b.shape[0]), n)
    d = (a[:real] - b[:real]).abs().max().item() if a.shape[0] >= real else float("nan")
    if d != 0:
        bad += 1
        if bad <= 12: print("FWD DIFF %-40s rows %d max abs %.2e  (first bad row %s)" % (name, real, d, int(((a[:real]-b[:real]).abs().amax(dim=tuple(range(1,a.dim()))) > 0).nonzero()[0])))
print("modules with forward differences:", bad, "of", len(acts))
# direct op probes at the deepest strided block
from d3feat.pytorch_b200 import ops
xe = acts["encoder_blocks.10"][0]; xs_ = acts["encoder_blocks.10"][2]
print("block10 out equal on real rows:", bool(torch.equal(xe[:sizes[3]], xs_[:sizes[3]])))
mp_e = ops.max_pool(xe, exact["pools"][3]); mp_s = ops.max_pool(xs_, static["pools"][3])
print("max_pool L3->4 equal:", bool(torch.equal(mp_e[:sizes[4]], mp_s[:sizes[4]])), "max abs", float((mp_e[:sizes[4]]-mp_s[:sizes[4]]).abs().max()))
pe, ps = exact["pools"][3], static["pools"][3]
print("pools[3] exact shape", tuple(pe.shape), "static", tuple(ps.shape), "exact max", int(pe.max()), "static real-row max", int(ps[:sizes[4]].max()))
def cmp(x, y, name):
    w = sorted(((float((x[k]-y[k]).abs().max()/max(float(y[k].abs().max()),1e-30)), k) for k in x), reverse=True)[:3]
    print(name, ["%.1e %s" % t for t in w])
cmp(g_a2, g_a, "drop-in run-to-run"); cmp(g_b2, g_b, "static run-to-run"); cmp(g_b, g_a, "static vs drop-in")
