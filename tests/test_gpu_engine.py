"""The static-capacity / CUDA-graph pipeline (engine.PairStep) must reproduce the exact-shape drop-in
pipeline: identical pyramid on the real rows, same descriptors / loss / gradients."""
import numpy as np
import pytest
import torch

import _inputs
from _util import rel_err
from d3feat.pytorch_b200 import synthetic
from d3feat.pytorch_b200.config import build_architecture, default_config

pytestmark = pytest.mark.gpu


def _setup(cuda, deform=False):
    from d3feat.pytorch_b200.architectures import KPFCNN
    kw = dict(architecture=build_architecture(5, deformable_from=3)) if deform else {}
    cfg = default_config(first_features_dim=32, num_node=64, **kw)
    limits = [40, 40, 40, 120, 120] if deform else [40, 40, 40, 40, 40]
    model = KPFCNN(cfg).to(cuda)
    model.load_state_dict(_inputs.kpfcnn_state_dict(cfg, seed=3), strict=True)
    model.train()
    return cfg, limits, model


def test_collate_static_matches_exact_pyramid(cuda):
    from d3feat.pytorch_b200.dataloader import collate_fn_descriptor
    from d3feat.pytorch_b200.engine import collate_static, plan_capacities
    cfg, limits, _ = _setup(cuda)
    data = synthetic.fragment_pair(1500, seed=5, num_node=64)
    exact = collate_fn_descriptor([data], cfg, limits)
    sizes = [int(p.shape[0]) for p in exact["points"]]
    caps = plan_capacities([sizes], margin=1.25, align=32)
    dev = [torch.as_tensor(a).to(cuda) for a in data]
    batch, pyramid = collate_static(*dev, cfg, limits, caps)
    status = pyramid.join()
    assert int(status.abs().sum()) == 0
    for l, n in enumerate(sizes):
        assert batch["points"][l].shape[0] == caps[l]
        assert torch.equal(batch["points"][l][:n], exact["points"][l])
        for key in ("neighbors", "pools", "upsamples"):
            e, s = exact[key][l], batch[key][l]
            if e.numel() == 0:
                continue
            rows, n_sup = e.shape[0], (sizes[l] if key != "upsamples" else sizes[l + 1])
            cap_sup = caps[l] if key != "upsamples" else caps[l + 1]
            assert s.shape[1] == limits[l] and s.dtype == torch.int32
            w = e.shape[1]
            got = s[:rows, :w].long()
            want = torch.where(e == n_sup, torch.full_like(e, cap_sup), e)   # shadow index = capacity
            assert torch.equal(got, want), (key, l)
            assert bool((s[:rows, w:] == cap_sup).all()) and bool((s[rows:] == cap_sup).all())
        assert batch["stack_lengths"][l].tolist() == exact["stack_lengths"][l].tolist()


def test_collate_static_side_streams_same_pyramid(cuda):
    """Subsampling chain and searches on side streams + transposed lists: same tensors as the in-order build."""
    from d3feat.pytorch_b200.engine import collate_static, plan_capacities
    cfg, limits, _ = _setup(cuda)
    data = synthetic.fragment_pair(1500, seed=6, num_node=64)
    dev = [torch.as_tensor(a).to(cuda) for a in data]
    ref, pyr = collate_static(*dev, cfg, limits, [3000, 1280, 384, 128, 64])
    pyr.join()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    got, pyr2 = collate_static(*dev, cfg, limits, [3000, 1280, 384, 128, 64], side_stream=s1, transposes=True, search_stream=s2,
                               transpose_stream=torch.cuda.Stream())
    status = pyr2.join()
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0
    for key in ("points", "neighbors", "pools", "upsamples", "stack_lengths"):
        for a, b in zip(ref[key], got[key]):
            assert torch.equal(a, b), key
    for l, m in enumerate(got["neighbors"]):
        t_off, t_src = m._d3f_transpose
        ns = got["points"][l].shape[0]
        assert int(t_off[-1]) == int((m < ns).sum()) and int(t_off[0]) == 0


@pytest.mark.parametrize("deform", [False, True])
@pytest.mark.parametrize("graph", [False, True])
def test_pair_step_matches_drop_in_pipeline(cuda, deform, graph):
    from d3feat.pytorch_b200.blocks import gather
    from d3feat.pytorch_b200.dataloader import collate_fn_descriptor
    from d3feat.pytorch_b200.engine import PairStep, plan_capacities
    from d3feat.pytorch_b200.loss import PairLoss
    cfg, limits, model = _setup(cuda, deform)
    loss_fn = PairLoss("circle", "euclidean", 10, 0.1, 0.1, 1.4)
    pairs = [synthetic.fragment_pair(1500, seed=5 + i, num_node=64) for i in range(2)]
    # reference: exact-shape drop-in pipeline, per-pair gradients
    def drop_in(data):
        batch = collate_fn_descriptor([data], cfg, limits)
        feats, scores = model(batch)
        c = batch["corr"].long()
        ia, ip = c[:, 0], c[:, 1] + 1500
        o = loss_fn(gather(feats, ia), gather(feats, ip), batch["dist_keypts"], gather(scores, ia), gather(scores, ip))
        model.zero_grad(set_to_none=True)
        (o["desc_loss"] + o["det_loss"]).backward()
        return ({k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None},
                float(o["desc_loss"].detach()), float(o["det_loss"].detach()))
    grads = [drop_in(d) for d in pairs]
    sizes = [[int(p.shape[0]) for p in collate_fn_descriptor([d], cfg, limits)["points"]] for d in pairs]
    caps = plan_capacities(sizes, margin=1.2, align=32)

    class _NoOpt:  # gradients only: keep the weights fixed so both pipelines see the same parameters
        def zero_grad(self, set_to_none=False):
            for p in model.parameters():
                if p.grad is not None:
                    p.grad.zero_()
        def step(self):
            pass
    for p in model.parameters():
        p.grad = torch.zeros_like(p) if p.requires_grad else None
    st = PairStep(model, cfg, limits, caps, 1500, 1500, loss_fn, _NoOpt(), None, num_node=64)
    if graph:
        st.capture()
    for data, (g_ref, dl, det) in zip(pairs, grads):
        st(data)
        st.check()
        assert abs(float(st.desc_loss) - dl) < 1e-5 * max(1.0, abs(dl)) and abs(float(st.det_loss) - det) < 1e-5 * max(1.0, abs(det))
        for k, p in model.named_parameters():
            if k in g_ref:
                assert rel_err(p.grad.cpu(), g_ref[k].cpu()) < 1e-4, k


def test_pair_step_reports_capacity_overflow(cuda):
    from d3feat.pytorch_b200.engine import PairStep
    from d3feat.pytorch_b200.loss import PairLoss
    cfg, limits, model = _setup(cuda)
    st = PairStep(model, cfg, limits, [3000, 64, 64, 64, 64], 1500, 1500, PairLoss("circle"), None, None, num_node=64)
    st(synthetic.fragment_pair(1500, seed=5, num_node=64))
    with pytest.raises(RuntimeError, match="overflow"):
        st.check()
