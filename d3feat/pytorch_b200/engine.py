"""Static-shape, sync-free execution of the whole hot path for one fragment pair, and its CUDA-graph
capture.

The drop-in API (dataloader.collate_fn_descriptor + KPFCNN + losses) sizes every tensor from data
(numbers of sub-sampled points, neighbour-matrix widths), which costs ~17 host round trips per pair
and ~700 eagerly launched kernels: on a B200 the step is then bound by the host, not the GPU.  This
module runs the same kernels on CAPACITY-padded tensors instead:

* level l of the pyramid is allocated with `caps[l]` rows (caps[0] = the real stacked size); the real
  row counts live on the device (the `lengths` vectors the C ABI already takes);
* padding query rows get all-shadow neighbour rows, padding support rows are never referenced, and
  the shadow index of level l is `caps[l]` (d3f_radius_neighbors' `pad_index`), so every real row
  computes exactly what it computes in the exact-shape pipeline;
* nothing reads back to the host inside the step, so collate + forward + loss + backward + optimizer
  is ONE `torch.cuda.CUDAGraph` replayed per pair; the overflow/validity flags of all kernels come back
  in one small status tensor next to the loss.

Neighbour matrices are int32 here (the kernels take either width).
"""
import math

import torch

from . import _lib, ops
from .blocks import gather


def plan_capacities(level_sizes, margin=1.10, align=64):
    """caps[l] for l >= 1 from observed level sizes (list of per-pair lists); level 0 is exact."""
    n_levels = len(level_sizes[0])
    caps = [max(s[0] for s in level_sizes)]
    for l in range(1, n_levels):
        m = max(s[l] for s in level_sizes)
        caps.append(int(math.ceil(m * margin / align) * align))
    return caps


def _pyramid_levels(config):
    """Per pyramid level: (has_conv, conv_is_deformable, has_pool, pool_is_deformable) -- the block walk of
    dataloader.py:95-170."""
    arch = config.architecture
    levels, layer_blocks = [], []
    for bi, block in enumerate(arch):
        if 'global' in block or 'upsample' in block:
            break
        if not ('pool' in block or 'strided' in block):
            layer_blocks.append(block)
            if bi < len(arch) - 1 and 'upsample' not in arch[bi + 1]:
                continue
        has_pool = 'pool' in block or 'strided' in block
        levels.append((bool(layer_blocks), any('deformable' in b for b in layer_blocks[:-1]), has_pool,
                       has_pool and 'deformable' in block))
        layer_blocks = []
    return levels


class _Pyramid:
    """Hand-over between the pyramid build (side streams) and its consumers (the model on the current stream)."""

    def __init__(self, main, streams):
        self.main, self.streams, self.events, self.flags = main, [st for st in streams if st is not None], {}, []

    def mark(self, key, stream):
        if stream is not None:
            ev = torch.cuda.Event()
            ev.record(stream)
            self.events[key] = ev

    def wait(self, *keys):
        """Make the CURRENT stream wait until the tensors named by `keys` have been produced."""
        cur = torch.cuda.current_stream()
        for key in keys:
            ev = self.events.get(key)
            if ev is not None:
                cur.wait_event(ev)

    def join(self):
        """Close the fork (every side stream back into the main stream) and return the status vector."""
        for st in self.streams:
            self.main.wait_stream(st)
        return torch.cat(self.flags)


def collate_static(pts0, pts1, feat0, feat1, corr, dist_keypts, config, limits, caps, lengths=None, side_stream=None,
                   transposes=False, search_stream=None, transpose_stream=None):
    """Same pyramid as dataloader.collate_fn_descriptor (reference dataloader.py:69-189) on capacity-padded tensors,
    with no host synchronisation.  Returns (batch dict, pyramid); `pyramid.join()` gives the status int32 tensor, which
    must be all zeros for the batch to be valid (checked by the caller after the step).

    Streams.  The grid-subsampling chain (level l+1 needs level l only; its order kernel is one CTA per cloud, ~130 us
    per level with 146 SMs idle) runs on `side_stream`, the radius searches on `search_stream`, the transposed lists
    (read by the backward pass only: with them on the search stream the first convolution started 140 us late) on
    `transpose_stream`, each behind the event of its search, and the CONSUMER -- the network on the current stream -- waits per tensor (batch['_pyramid'].wait(('neighbors', l))
    in blocks._conv_geometry etc.), so the level-0 convolutions start as soon as the level-0 search is done while the
    deeper levels are still being built.  Inside a CUDA-graph capture this becomes a fork/join in the graph.  With both
    streams None everything runs in order on the current stream."""
    dev = pts0.device
    points = torch.cat([pts0, pts1], dim=0)
    feats = torch.cat([feat0, feat1], dim=0)
    if lengths is None:  # (inside a CUDA-graph capture the caller passes a pre-built device tensor)
        lengths = torch.tensor([pts0.shape[0], pts1.shape[0]], dtype=torch.int32, device=dev)
    levels = _pyramid_levels(config)
    out = {'points': [], 'neighbors': [], 'pools': [], 'upsamples': [], 'stack_lengths': []}
    empty_idx = torch.zeros((0, 1), dtype=torch.int32, device=dev)

    main = torch.cuda.current_stream()
    pyr = _Pyramid(main, [side_stream, search_stream, transpose_stream])
    flags = pyr.flags
    for st in pyr.streams:
        st.wait_stream(main)             # fork: the inputs were written on the main stream

    # ---- subsampling chain
    sub_s = side_stream if side_stream is not None else main
    pts, lens = [points], [lengths]
    with torch.cuda.stream(sub_s):
        r_normal = config.first_subsampling_dl * config.conv_radius
        for l, (_, _, has_pool, _) in enumerate(levels):
            if has_pool:
                dl = 2 * r_normal / config.conv_radius
                pool_p, pool_len = ops.grid_subsample_raw(pts[l], lens[l], dl, caps[l + 1])
                flags.append(pool_len[2:3])                       # output capacity exceeded
                flags.append((pool_len[:2] < 0).to(torch.int32))  # unsupported voxel grid
                pts.append(pool_p)
                lens.append(pool_len[:2])
                pyr.mark(('points', l + 1), side_stream)
            else:
                pts.append(torch.zeros((0, 3), dtype=torch.float32, device=dev))
                lens.append(torch.zeros((0,), dtype=torch.int32, device=dev))
            r_normal *= 2

    def search(q, s, ql, sl, r, limit, pad, transpose=False):
        idx, info = ops.radius_neighbors_raw(q, s, ql, sl, r, int(limit), torch.int32, None, False, pad_index=pad)
        if transpose:   # training: "which queries list support j", consumed by every KPConv backward over this matrix
            if transpose_stream is not None:
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream())
                with torch.cuda.stream(transpose_stream):
                    transpose_stream.wait_event(done)
                    idx._d3f_transpose = ops.neighbors_transpose(idx, s.shape[0])
            else:
                idx._d3f_transpose = ops.neighbors_transpose(idx, s.shape[0])
        flags.append(info[1:2])           # 1 = a row overflowed the candidate buffer
        # the reference's matrix has min(max_count, limit) columns (dataloader.py:64-65): a row that fills it holds
        # no shadow index, which max_pool / the eval-mode detection gate can see -> keep that width on the device
        idx._d3f_width = torch.clamp(info[0:1], max=int(limit))
        return idx

    # ---- radius searches
    srch_s = search_stream if search_stream is not None else main
    with torch.cuda.stream(srch_s):
        r_normal = config.first_subsampling_dl * config.conv_radius
        deferred = []
        for l, (has_conv, deform_conv, has_pool, deform_pool) in enumerate(levels):
            cap = caps[l]
            pyr.wait(('points', l))                      # (no-op for level 0 and without a side stream)
            if has_conv:
                r = r_normal * config.deform_radius / config.conv_radius if deform_conv else r_normal
                conv_i = search(pts[l], pts[l], lens[l], lens[l], r, limits[l], cap, transposes)
                pyr.mark(('neighbors', l), search_stream)
            else:
                conv_i = empty_idx
            if has_pool:
                pyr.wait(('points', l + 1))
                r = r_normal * config.deform_radius / config.conv_radius if deform_pool else r_normal
                pool_i = search(pts[l + 1], pts[l], lens[l + 1], lens[l], r, limits[l], cap, transposes)
                pyr.mark(('pools', l), search_stream)
                up_i = None     # only the decoder reads the upsampling matrices: searched after every encoder matrix
                deferred.append((l, 2 * r))
            else:
                pool_i, up_i = empty_idx, empty_idx
            out['points'].append(pts[l])
            out['neighbors'].append(conv_i)
            out['pools'].append(pool_i)
            out['upsamples'].append(up_i)
            out['stack_lengths'].append(lens[l])
            r_normal *= 2
        for l, r_up in deferred:
            out['upsamples'][l] = search(pts[l], pts[l + 1], lens[l], lens[l + 1], r_up, limits[l], caps[l + 1])
            pyr.mark(('upsamples', l), search_stream)
    if search_stream is None and side_stream is not None:
        main.wait_stream(side_stream)    # searches ran on the main stream: everything the consumer needs is ordered
    out['features'] = feats
    out['corr'] = corr
    out['dist_keypts'] = dist_keypts
    out['_pyramid'] = pyr
    return out, pyr


class PairStep:
    """collate -> KPFCNN -> descriptor + detector loss (-> backward -> optimizer) for pairs of a fixed size,
    eagerly on static shapes or as one CUDA graph.

    step = PairStep(model, config, limits, caps, n0, n1, loss_fn, optimizer, flat_grads)
    step.capture()                      # optional: CUDA graph
    loss = step(data)                   # data = (pts0, pts1, feat0, feat1, corr, dist_keypts): tensors on any device
    step.check()                        # raises if ANY step since the last check overflowed a capacity / candidate
                                        # buffer, hit a GEMM barrier timeout, or was skipped for a non-finite gradient

    `optimizer`: optim.FlatSGD (the reference's SGD + ExpLR + non-finite guard on flat buffers, gradients written in
    place, bucketed all-reduce overlapped with backward) or any object with zero_grad() / step() (torch.optim).
    """

    STATUS_EXTRA = 2     # [gemm barrier timeout, step skipped: non-finite gradient] after the pyramid flags

    def __init__(self, model, config, limits, caps, n0, n1, loss_fn, optimizer=None, flat_grads=None,
                 num_node=128, group=None, cross_fragment=None):
        from .optim import FlatSGD, early_block_index
        dev = next(model.parameters()).device
        self.model, self.config, self.limits, self.caps = model, config, [int(v) for v in limits], list(caps)
        self.loss_fn, self.optimizer, self.flat = loss_fn, optimizer, flat_grads
        self.cross_fragment, self.group = cross_fragment, group
        self.flat_sgd = optimizer if isinstance(optimizer, FlatSGD) else None
        self.n0 = n0
        f32 = dict(dtype=torch.float32, device=dev)
        self.inputs = (torch.zeros((n0, 3), **f32), torch.zeros((n1, 3), **f32), torch.ones((n0, 1), **f32),
                       torch.ones((n1, 1), **f32), torch.zeros((num_node, 2), dtype=torch.int64, device=dev),
                       torch.zeros((num_node, num_node), dtype=torch.float64, device=dev))
        self.lengths0 = torch.tensor([n0, n1], dtype=torch.int32, device=dev)
        self.graph = None
        self.zero_arena = None
        self.loss = torch.zeros((), **f32)
        self.desc_loss = torch.zeros((), **f32)
        self.det_loss = torch.zeros((), **f32)
        self.status = None          # sticky: max over every step since the last check()
        self.batch = None
        self.features = None        # [caps[0], 32] descriptors / [caps[0], 1] scores of the last step (static buffers)
        self.scores = None
        self.side_stream = torch.cuda.Stream(device=dev)     # grid-subsampling chain
        self.transpose_stream = torch.cuda.Stream(device=dev)   # transposed neighbour lists (backward only)
        self.search_stream = torch.cuda.Stream(device=dev)   # radius searches; the network consumes
                                                             # each level as soon as it is ready (see collate_static)
        if self.flat_sgd is not None and cross_fragment is not None and hasattr(model, "_early_block"):
            # data parallel: all-reduce the first gradient bucket while the shallow levels still back-propagate
            model._early_block = early_block_index(model)
            model._on_early_grads = lambda: self.flat_sgd.allreduce_early(self.group)

    # -- the step on the static input buffers
    def _body(self):
        # without an optimizer the step is forward + loss only: no autograd graph is recorded
        from . import ops
        if self.zero_arena is None:
            self.zero_arena = ops.ZeroArena()
        prev, ops.ZERO_ARENA = ops.ZERO_ARENA, self.zero_arena
        try:
            with torch.set_grad_enabled(self.optimizer is not None):
                self.zero_arena.reset()       # one fill for every split-K GEMM output of the step
                return self._body_impl()
        finally:
            ops.ZERO_ARENA = prev

    def _body_impl(self):
        cfg = self.config
        lib = _lib.load()
        if self.flat_sgd is not None:
            self.flat_sgd.zero_grad()     # one fill, while the main stream would otherwise wait for the first search
        batch, pyramid = collate_static(*self.inputs, cfg, self.limits, self.caps, self.lengths0, self.side_stream,
                                        transposes=self.optimizer is not None, search_stream=self.search_stream,
                                        transpose_stream=self.transpose_stream)
        feats, scores = self.model(batch)
        status = pyramid.join()          # the backward pass reads the transposed lists built on the search stream
        c = batch['corr']
        ia, ip = c[:, 0], c[:, 1] + self.n0
        a, p = gather(feats, ia), gather(feats, ip)            # trainer.py:91-94
        sa, sp = gather(scores, ia), gather(scores, ip)
        if self.cross_fragment is not None:
            out = self.cross_fragment(self.loss_fn, a, p, batch['dist_keypts'], sa, sp, self.group)
        else:
            out = self.loss_fn(a, p, batch['dist_keypts'], sa, sp)
        loss = out['desc_loss'] * cfg.desc_loss_weight + out['det_loss'] * cfg.det_loss_weight
        extra = torch.zeros(self.STATUS_EXTRA, dtype=torch.int32, device=status.device)
        if self.optimizer is not None:
            if self.flat is not None:
                self.flat.zero()
            else:
                # FlatSGD(direct): every gradient is written in place by the kernel that computes it (the flat buffer was
                # cleared at the top of the step).  torch.optim: gradients are (re)created by the backward pass (inside a
                # CUDA graph they come from the graph's private pool at the same addresses every replay).  Either way:
                # no per-parameter zero-fill, no accumulate kernel per parameter
                if self.flat_sgd is None:
                    self.optimizer.zero_grad(set_to_none=True)
            loss.backward()
            if self.flat_sgd is not None:
                self.flat_sgd.allreduce(self.group)
            elif self.flat is not None:
                self.flat.allreduce(self.group)
            self.optimizer.step()       # FlatSGD: skipped on the device when a gradient is not finite (trainer.py:104-111)
            if self.flat_sgd is not None:
                extra[1:2].copy_(self.flat_sgd.nonfinite)
        _lib.check(lib.d3f_gemm_status_snapshot(extra.data_ptr(), torch.cuda.current_stream().cuda_stream))
        status = torch.cat([status, extra])
        self.loss.copy_(loss.detach())
        self.desc_loss.copy_(out['desc_loss'].detach())
        self.det_loss.copy_(out['det_loss'].detach())
        if self.status is None:
            self.status = torch.zeros_like(status)
        torch.maximum(self.status, status, out=self.status)      # sticky across steps / graph replays
        self.batch = batch
        self.features, self.scores = feats.detach(), scores.detach()
        return loss

    def load(self, data):
        for dst, src in zip(self.inputs, data):
            if not isinstance(src, torch.Tensor):
                src = torch.as_tensor(src)
            dst.copy_(src, non_blocking=True)

    def capture(self, warmup=3):
        """Warm up on a side stream (lazy optimizer state, cudaFuncSetAttribute calls, allocator), then capture."""
        import gc
        gc.collect()   # drop autograd graphs (and their AccumulateGrad nodes) left over from eager steps on other streams
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        # deformable KPConv modules keep graph tensors of the last forward (deformed_KP, offset_features: the
        # reference API for its regulariser); left alive they pin the warm-up iteration's autograd graph and
        # its AccumulateGrad nodes, which belong to another stream and would invalidate the capture
        for m in self.model.modules():
            if hasattr(m, "deformed_KP"):
                m.deformed_KP = m.offset_features = m.min_d2 = None
        self.batch = self.features = self.scores = None
        gc.collect()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._body()
        self.graph = g
        return self

    def release(self):
        """Drop the captured graph (and the NCCL kernels it holds) -- call before destroying the process group."""
        self.graph = None
        self.batch = self.features = self.scores = None
        if hasattr(self.model, "_on_early_grads"):
            self.model._on_early_grads = None

    def __call__(self, data=None):
        if data is not None:
            self.load(data)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        return self.loss

    def check(self, allow_skipped=False):
        """Host-side validity check of EVERY step since the previous check (one small D2H read; clears the flags).
        allow_skipped: a step skipped by the non-finite-gradient guard is what the reference does silently
        (trainer.py:104-111); a benchmark must not tolerate it (skipped work), a training loop may."""
        if self.status is None:
            return
        st = self.status.tolist()
        self.status.zero_()
        if allow_skipped:
            st[-1] = 0
        if any(st):
            pyr, (gemm_to, skipped) = st[:-self.STATUS_EXTRA], st[-self.STATUS_EXTRA:]
            what = []
            if any(pyr):
                what.append("a static capacity or a neighbour candidate buffer overflowed (pyramid flags %s); re-plan the "
                            "capacities or use the exact-shape pipeline" % pyr)
            if gemm_to:
                what.append("a tcgen05 GEMM timed out on its mbarrier (output tile NaN-poisoned)")
            if skipped:
                what.append("an optimizer step was skipped: non-finite gradient (trainer.py:104-111 semantics)")
            raise RuntimeError("PairStep: " + "; ".join(what))
