import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import _inputs
from d3feat.pytorch_b200 import synthetic
from d3feat.pytorch_b200.config import default_config
from d3feat.pytorch_b200.architectures import KPFCNN
from d3feat.pytorch_b200.dataloader import collate_fn_descriptor
from oracle import model_ref, pipeline
cfg = default_config(); limits = [35, 42, 42, 45, 47]
for seed in (0, 1):
    data = synthetic.fragment_pair(20000, seed=seed)
    sd = _inputs.kpfcnn_state_dict(cfg, seed=0)
    cpu_b = pipeline.cpu_collate(data, cfg, limits, impl="ref")
    with torch.no_grad():
        f_ref, s_ref = model_ref.kpfcnn_forward(sd, cpu_b, cfg, training=True)
    batch = collate_fn_descriptor([data], cfg, limits)
    model = KPFCNN(cfg).cuda(); model.load_state_dict(sd); model.train()
    with torch.no_grad():
        f, s = model(batch)
    d = (s.cpu() - s_ref).abs().reshape(-1)
    print("seed", seed, "max ref score", float(s_ref.abs().max()), "max diff", float(d.max()), "n > 1e-5", int((d > 1e-5).sum()),
          "argmax", int(d.argmax()), "ref", float(s_ref.reshape(-1)[d.argmax()]), "gpu", float(s.reshape(-1)[d.argmax()]))
    i = int(d.argmax())
    nb = cpu_b["neighbors"][0][i]
    print("  neighbours of worst row: count real", int((nb < 40000).sum()), "width", nb.shape[0])
    top = torch.topk(d, 5)
    print("  top diffs", top.values.tolist(), top.indices.tolist())
    if seed == 0:
        bad = torch.nonzero(d > 1e-5).reshape(-1)
        nbs = cpu_b["neighbors"][0][bad]
        vals, cnts = torch.unique(nbs[nbs < 40000], return_counts=True)
        common = vals[cnts == cnts.max()]
        print("  bad rows", bad.tolist()[:40])
        print("  most common neighbours", common.tolist(), "count", int(cnts.max()), "of", len(bad))
        # un-normalised descriptors: recompute
        feats = {}
        def run(collect):
            return model_ref.kpfcnn_forward(sd, cpu_b, cfg, training=True, collect=collect)
        import inspect
        print(inspect.signature(model_ref.kpfcnn_forward), inspect.getsource(model_ref.detection_scores)[:1500])
