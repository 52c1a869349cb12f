"""Tensor-level wrappers over the C ABI (include/d3feat_b200.h).

torch is used for device memory and streams only: every function allocates its outputs
and workspace with torch, passes raw device pointers plus the current CUDA stream to
libd3feat_b200.so, and returns torch tensors.  No function here has a CPU path.
"""
import torch

from . import _lib

INFLUENCE = {"constant": 0, "linear": 1, "gaussian": 2}
AGGREGATION = {"sum": 0, "closest": 1}
METRIC = {"euclidean": 0, "sqeuclidean": 1, "cityblock": 2, "cosine": 3, "arccosine": 4}
LOSS_KIND = {"circle": 0, "contrastive": 1}

launch_count = 0  # number of C-ABI hot-path calls issued

# Optional per-op CUDA-event timing (bench.py switches it on for its timed region): a dict
# tag -> [(start_event, end_event), ...] recorded on the stream the kernels are launched on.
PROFILE = None


class _Timed:
    __slots__ = ("tag", "ev")

    def __init__(self, tag):
        self.tag = tag
        self.ev = None

    def __enter__(self):
        if PROFILE is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            self.ev[1].record()
            PROFILE.setdefault(self.tag, []).append(self.ev)
        return False


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cuda_f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("d3feat.pytorch_b200: `%s` must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _pow2ceil(v):
    p = 1
    while p < v:
        p <<= 1
    return p


# --------------------------------------------------------------------------- radius neighbours
def radius_neighbors_raw(queries, supports, q_len, s_len, radius, max_cols, index_dtype=torch.int64,
                         row_capacity=None, count_only=False, pad_index=-1):
    """One call of d3f_radius_neighbors (no host sync).  Returns (idx [Nq,max_cols] or None, info int32[4] on
    device).  `queries` / `supports` may have more rows than sum(q_len) / sum(s_len) (static capacities)."""
    lib = _lib.load()
    queries, supports = _cuda_f32(queries, "queries"), _cuda_f32(supports, "supports")
    dev = queries.device
    q_len = q_len.to(device=dev, dtype=torch.int32).contiguous()
    s_len = s_len.to(device=dev, dtype=torch.int32).contiguous()
    nq, ns, nb = queries.shape[0], supports.shape[0], q_len.shape[0]
    info = torch.empty(4, dtype=torch.int32, device=dev)
    if row_capacity is None:
        row_capacity = min(8192, max(64, _pow2ceil(2 * max(int(max_cols), 1))))
    out = None
    if not count_only:
        out = torch.empty((nq, int(max_cols)), dtype=index_dtype, device=dev)
    ws_bytes = lib.d3f_radius_neighbors_workspace_bytes(nq, ns, nb)
    ws = _ws(ws_bytes, dev)
    global launch_count
    launch_count += 1
    with _Timed(("radius_neighbors", nq, ns)):
      _lib.check(lib.d3f_radius_neighbors(_p(queries), _p(supports), _p(q_len), _p(s_len), nb, nq, ns,
                                        float(radius), int(max_cols) if not count_only else 0, _p(out),
                                        1 if index_dtype == torch.int64 else 0, int(pad_index), _p(info), int(row_capacity),
                                        _p(ws), ws.numel(), _stream()))
    return out, info


def radius_neighbors(queries, supports, q_len, s_len, radius, max_neighbors=0, index_dtype=torch.int64):
    """Drop-in semantics of batch_neighbors_kpconv (datasets/dataloader.py:52-67): the matrix has
    min(max_count, max_neighbors) columns (all of them if max_neighbors <= 0).  Reads 2 ints back
    from the device to learn the width, exactly the information the reference gets from the shape
    of the NumPy array."""
    if max_neighbors <= 0:
        _, info = radius_neighbors_raw(queries, supports, q_len, s_len, radius, 0, index_dtype, 64, count_only=True)
        max_neighbors = max(int(info[0].item()), 1)
    cap = None
    while True:
        out, info = radius_neighbors_raw(queries, supports, q_len, s_len, radius, max_neighbors, index_dtype, cap)
        h = info[:2].tolist()
        if h[1] == 0:
            break
        if h[0] > 8192:
            raise _lib.D3FError("a query has %d in-range supports; more than the 8192-entry row buffer" % h[0])
        cap = _pow2ceil(h[0])  # a row overflowed the candidate buffer: redo with room for the largest row
    width = min(h[0], int(max_neighbors))
    return out[:, :width].contiguous() if width < out.shape[1] else out


# --------------------------------------------------------------------------- grid subsampling
def grid_subsample(points, lengths, sample_dl):
    """Drop-in semantics of batch_grid_subsampling_kpconv (datasets/dataloader.py:12-21), points only.
    Returns (s_points [M,3] f32, s_len [B] i32) on the device; reads B ints back to size s_points."""
    lib = _lib.load()
    points = _cuda_f32(points, "points")
    dev = points.device
    lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
    out, out_len = grid_subsample_raw(points, lengths, sample_dl, points.shape[0])
    host_len = out_len.tolist()
    if any(v < 0 for v in host_len[:-1]):
        raise _lib.D3FError("grid_subsample: voxel grid exceeds the supported key range (code %s)" % host_len)
    return out[:sum(host_len[:-1])], out_len[:-1]


def grid_subsample_raw(points, lengths, sample_dl, out_capacity):
    """One call of d3f_grid_subsample (no host sync) -> (out [out_capacity,3], out_len int32 [B+1] on device;
    out_len[B] = 1 if out_capacity was too small, out_len[b] = -1 on an unsupported voxel grid)."""
    lib = _lib.load()
    points = _cuda_f32(points, "points")
    dev = points.device
    lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
    n, nb = points.shape[0], lengths.shape[0]
    out = torch.empty((int(out_capacity), 3), dtype=torch.float32, device=dev)
    out_len = torch.empty(nb + 1, dtype=torch.int32, device=dev)
    ws = _ws(lib.d3f_grid_subsample_workspace_bytes(n, nb), dev)
    global launch_count
    launch_count += 1
    with _Timed(("grid_subsample", n)):
      _lib.check(lib.d3f_grid_subsample(_p(points), _p(lengths), nb, n, float(sample_dl), _p(out), int(out_capacity),
                                      _p(out_len), _p(ws), ws.numel(), _stream()))
    return out, out_len


# --------------------------------------------------------------------------- KPConv
def kpconv_fused_eligible(H, K, cin, cout, deformed=False, modulations=None, influence="linear", aggregation="sum"):
    """True if d3f_kpconv_forward_ex runs this layer as ONE fused kernel (then `wf` is optional)."""
    return (not deformed and modulations is None and influence == "linear" and aggregation == "sum"
            and bool(_lib.load().d3f_kpconv_fused_eligible(int(H), int(K), int(cin), int(cout))))


def kpconv_forward(q_pts, s_pts, inds, x, weights, kernel_points, extent, influence, aggregation,
                   deformed=False, modulations=None, want_min_d2=False, bias=None, slope=None, need_wf=True):
    """d3f_kpconv_forward_ex: out = act(KPConv(...) + bias) with act = LeakyReLU(slope) when slope is given.
    need_wf=False: the caller does not need the kernel-point-weighted features [Nq, K, Cin] (inference, or a backward
    that works from the transposed neighbour lists); on the fused path they are then never written (returned as None),
    on the two-kernel path they remain the contraction's input."""
    lib = _lib.load()
    q_pts, s_pts, x = _cuda_f32(q_pts, "q_pts"), _cuda_f32(s_pts, "s_pts"), _cuda_f32(x, "x")
    weights, kernel_points = _cuda_f32(weights, "weights"), _cuda_f32(kernel_points, "kernel_points")
    if not inds.is_cuda:
        raise RuntimeError("d3feat.pytorch_b200: `neighb_inds` must be a CUDA tensor")
    if inds.dtype not in (torch.int32, torch.int64):
        inds = inds.long()
    if inds.stride(-1) != 1:
        inds = inds.contiguous()
    dev = x.device
    nq, ns, H = q_pts.shape[0], s_pts.shape[0], inds.shape[1]
    K, cin, cout = weights.shape
    if x.shape[0] != ns or x.shape[1] != cin:
        raise RuntimeError("KPConv: x must be [n_supports, in_channels] = [%d, %d], got %s" % (ns, cin, tuple(x.shape)))
    out = torch.empty((nq, cout), dtype=torch.float32, device=dev)
    skip_wf = (not need_wf) and ns > 0 and nq > 0 and kpconv_fused_eligible(H, K, cin, cout, deformed, modulations, influence,
                                                                            aggregation)
    wf = None if skip_wf else torch.empty((nq, K, cin), dtype=torch.float32, device=dev)
    wf_un = torch.empty_like(wf) if modulations is not None else None
    inv_n = torch.empty(nq, dtype=torch.float32, device=dev)
    min_d2 = torch.empty((nq, K), dtype=torch.float32, device=dev) if (deformed and want_min_d2) else None
    if modulations is not None:
        modulations = _cuda_f32(modulations, "modulations")
    if bias is not None:
        bias = _cuda_f32(bias, "bias")
        if bias.numel() != cout:
            raise RuntimeError("KPConv: bias must have out_channels = %d elements" % cout)
    ws = _ws(lib.d3f_kpconv_workspace_bytes(nq, ns, H, K, cin, cout), dev)
    global launch_count
    launch_count += 1
    gev = None
    if PROFILE is not None:   # time the gather kernel alone as well (events recorded inside the C call)
        gev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        gev[0].record(); gev[1].record()   # torch creates the cudaEvent_t lazily: record once to get the handles
        lib.d3f_kpconv_set_gather_events(gev[0].cuda_event, gev[1].cuda_event)
        PROFILE.setdefault(("kpconv_gather", nq, ns, H, cin, cout, bool(deformed)), []).append(gev)
    with _Timed(("kpconv_fwd", nq, ns, H, cin, cout, bool(deformed))):
      _lib.check(lib.d3f_kpconv_forward_ex(_p(q_pts), _p(s_pts), _p(inds), 1 if inds.dtype == torch.int64 else 0,
                                         inds.stride(0) if H > 0 else 0, _p(x), _p(weights), _p(kernel_points),
                                         1 if deformed else 0, _p(modulations), nq, ns, H, K, cin, cout,
                                         float(extent), INFLUENCE[influence], AGGREGATION[aggregation],
                                         _p(bias), 0 if slope is None else 1, 0.0 if slope is None else float(slope),
                                         _p(out), _p(wf), _p(wf_un), _p(inv_n), _p(min_d2), _p(ws), ws.numel(), _stream()))
    if gev is not None:
        lib.d3f_kpconv_set_gather_events(None, None)
    return out, wf, wf_un, inv_n, min_d2


def neighbors_transpose(inds, n_supports):
    """CSR lists "which queries reference support j" of a neighbour matrix [Nq, H] (d3f_neighbors_transpose):
    (t_offsets int32 [n_supports + 1], t_src int32 [Nq * H]).  Feeds the atomic-free KPConv backward."""
    lib = _lib.load()
    if not inds.is_cuda:
        raise RuntimeError("d3feat.pytorch_b200: `inds` must be a CUDA tensor")
    if inds.dtype not in (torch.int32, torch.int64):
        inds = inds.long()
    if inds.dim() != 2 or (inds.shape[1] > 0 and inds.stride(-1) != 1):
        inds = inds.contiguous()
    nq, H = inds.shape
    dev = inds.device
    t_off = torch.empty(int(n_supports) + 1, dtype=torch.int32, device=dev)
    t_src = torch.empty(max(nq * H, 1), dtype=torch.int32, device=dev)
    ws = _ws(lib.d3f_neighbors_transpose_workspace_bytes(int(n_supports)), dev)
    global launch_count
    launch_count += 1
    with _Timed(("neighbors_transpose", nq, int(n_supports), H)):
        _lib.check(lib.d3f_neighbors_transpose(_p(inds), 1 if inds.dtype == torch.int64 else 0, inds.stride(0) if H > 0 else 0,
                                               nq, int(n_supports), H, _p(t_off), _p(t_src), _p(ws), ws.numel(), _stream()))
    return t_off, t_src


def kpconv_backward(q_pts, s_pts, inds, x, weights, kernel_points, extent, influence, aggregation, deformed,
                    modulations, wf, wf_un, inv_n, grad_out, need_x, need_w, need_kp, need_mod, transpose=None, gw_out=None):
    """d3f_kpconv_backward_ex.  `transpose` = (t_offsets, t_src) from neighbors_transpose selects the atomic-free
    grad_x where the layer allows it (rigid, Cout % 32 == 0)."""
    lib = _lib.load()
    dev = x.device
    nq, ns, H = q_pts.shape[0], s_pts.shape[0], inds.shape[1]
    K, cin, cout = weights.shape
    grad_out = _cuda_f32(grad_out, "grad_out")
    gx = torch.empty((ns, cin), dtype=torch.float32, device=dev) if need_x else None
    gw = (gw_out if gw_out is not None else torch.empty_like(weights)) if need_w else None
    gkp = torch.empty((nq, K, 3), dtype=torch.float32, device=dev) if (need_kp and deformed) else None
    gmod = torch.empty((nq, K), dtype=torch.float32, device=dev) if (need_mod and modulations is not None) else None
    ws = _ws(lib.d3f_kpconv_workspace_bytes(nq, ns, H, K, cin, cout), dev)
    global launch_count
    launch_count += 1
    with _Timed(("kpconv_bwd", nq, ns, H, cin, cout, bool(deformed))):
      t_off, t_src = transpose if transpose is not None else (None, None)
      _lib.check(lib.d3f_kpconv_backward_ex(_p(q_pts), _p(s_pts), _p(inds), 1 if inds.dtype == torch.int64 else 0,
                                          inds.stride(0) if H > 0 else 0, _p(x), _p(weights), _p(kernel_points),
                                          1 if deformed else 0, _p(modulations), nq, ns, H, K, cin, cout,
                                          float(extent), INFLUENCE[influence], AGGREGATION[aggregation],
                                          _p(wf), _p(wf_un), _p(inv_n), _p(grad_out), _p(gx), _p(gw), _p(gkp), _p(gmod),
                                          _p(t_off), _p(t_src), _p(ws), ws.numel(), _stream()))
    return gx, gw, gkp, gmod


def kpconv_gather_transposed(q_pts, s_pts, transpose, grad_out, inv_n, kernel_points, cout, extent, influence, aggregation):
    """G [Ns, K, Cout]: the forward gather over the transposed neighbour lists, reading rows of grad_out * (1/n)."""
    lib = _lib.load()
    t_off, t_src = transpose
    nq, ns, K = q_pts.shape[0], s_pts.shape[0], kernel_points.shape[0]
    grad_out = _cuda_f32(grad_out, "grad_out")
    G = torch.empty((ns, K, cout), dtype=torch.float32, device=grad_out.device)
    global launch_count
    launch_count += 1
    with _Timed(("kpconv_bwd_gather", nq, ns, cout)):
        _lib.check(lib.d3f_kpconv_gather_transposed(_p(q_pts), _p(s_pts), _p(t_off), _p(t_src), _p(grad_out), _p(inv_n),
                                                    _p(kernel_points), nq, ns, K, cout, float(extent), INFLUENCE[influence],
                                                    AGGREGATION[aggregation], _p(G), _stream()))
    return G


def kpconv_grads_from_gathered(G, x, weights, need_x, need_w, gw_out=None):
    """(grad_x [Ns, Cin] = G W^T, grad_weights [K, Cin, Cout] = x^T G) from the gathered gradient G."""
    lib = _lib.load()
    ns, K, cout = G.shape
    cin = weights.shape[1]
    gx = torch.empty((ns, cin), dtype=torch.float32, device=G.device) if need_x else None
    gw = (gw_out if gw_out is not None else torch.empty_like(weights)) if need_w else None
    with _Timed(("kpconv_bwd_gemm", ns, cin, cout, bool(need_x), bool(need_w))):
        _lib.check(lib.d3f_kpconv_grads_from_gathered(_p(G), _p(x), _p(weights), ns, K, cin, cout, _p(gx), _p(gw),
                                                      1 if (need_w and _prezeroed(gw_out)) else 0, _stream()))
    return gx, gw


# --------------------------------------------------------------------------- descriptor distance / losses
def pair_dist(a, b, metric="euclidean"):
    lib = _lib.load()
    a, b = _cuda_f32(a, "a"), _cuda_f32(b, "b")
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    global launch_count
    launch_count += 1
    _lib.check(lib.d3f_pair_dist(_p(a), _p(b), a.shape[0], b.shape[0], a.shape[1], METRIC[metric], _p(out), _stream()))
    return out


def _keypts(dist_keypts, dev):
    if not dist_keypts.is_cuda:
        dist_keypts = dist_keypts.to(dev)
    if dist_keypts.dtype not in (torch.float32, torch.float64):
        dist_keypts = dist_keypts.double()
    return dist_keypts.contiguous()


def pair_loss_forward(anchor, positive, dist_keypts, anc_score, pos_score, kind, metric, safe_radius,
                      pos_margin, neg_margin, log_scale):
    lib = _lib.load()
    anchor, positive = _cuda_f32(anchor, "anchor"), _cuda_f32(positive, "positive")
    dev = anchor.device
    P, D = anchor.shape
    dk = _keypts(dist_keypts, dev)
    sa = _cuda_f32(anc_score.reshape(-1), "anc_score") if anc_score is not None else None
    sp = _cuda_f32(pos_score.reshape(-1), "pos_score") if pos_score is not None else None
    dists = torch.empty((P, P), dtype=torch.float32, device=dev)
    stats = torch.empty(8, dtype=torch.float32, device=dev)
    fp = torch.empty(P, dtype=torch.float32, device=dev)
    an = torch.empty(P, dtype=torch.float32, device=dev)
    aux = torch.empty(lib.d3f_pair_loss_aux_floats(P), dtype=torch.float32, device=dev)
    global launch_count
    launch_count += 1
    _lib.check(lib.d3f_pair_loss_forward(_p(anchor), _p(positive), P, D, _p(dk), 1 if dk.dtype == torch.float64 else 0,
                                         _p(sa), _p(sp), LOSS_KIND[kind], METRIC[metric], float(safe_radius),
                                         float(pos_margin), float(neg_margin), float(log_scale),
                                         _p(dists), _p(stats), _p(fp), _p(an), _p(aux), _stream()))
    return dists, stats, fp, an, aux, (anchor, positive, dk, sa, sp)


def pair_loss_backward(saved, kind, metric, safe_radius, pos_margin, neg_margin, log_scale, dists, aux, grad_losses):
    lib = _lib.load()
    anchor, positive, dk, sa, sp = saved
    P, D = anchor.shape
    ga, gp = torch.empty_like(anchor), torch.empty_like(positive)
    gsa = torch.empty_like(sa) if sa is not None else None
    gsp = torch.empty_like(sp) if sp is not None else None
    global launch_count
    launch_count += 1
    _lib.check(lib.d3f_pair_loss_backward(_p(anchor), _p(positive), P, D, _p(dk), 1 if dk.dtype == torch.float64 else 0,
                                          _p(sa), _p(sp), LOSS_KIND[kind], METRIC[metric], float(safe_radius),
                                          float(pos_margin), float(neg_margin), float(log_scale),
                                          _p(dists), _p(aux), _p(grad_losses.contiguous()), _p(ga), _p(gp), _p(gsa), _p(gsp),
                                          _stream()))
    return ga, gp, gsa, gsp


# --------------------------------------------------------------------------- mutual nearest-neighbour matching
def mutual_nn(source, target):
    """d3f_mutual_nn: (pairs [n,2] int32 ascending in source index, source_arg [Ns] int32, target_arg [Nt] int32) for
    two descriptor sets on the GPU.  One D2H read of the pair count sizes the result, like the NumPy array the
    reference builds (geometric_registration/common.py:5-21)."""
    lib = _lib.load()
    source, target = _cuda_f32(source, "source"), _cuda_f32(target, "target")
    if source.dim() != 2 or target.dim() != 2 or source.shape[1] != target.shape[1]:
        raise RuntimeError("mutual_nn: descriptors must be [N, D] and [M, D]")
    ns, nt, d = source.shape[0], target.shape[0], source.shape[1]
    dev = source.device
    sarg = torch.empty(ns, dtype=torch.int32, device=dev)
    targ = torch.empty(nt, dtype=torch.int32, device=dev)
    pairs = torch.empty((max(ns, 1), 2), dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    global launch_count
    launch_count += 1
    _lib.check(lib.d3f_mutual_nn(_p(source), _p(target), ns, nt, d, _p(sarg), _p(targ), _p(pairs), _p(count), _stream()))
    return pairs[:int(count.item())], sarg, targ


# --------------------------------------------------------------------------- pooling / gathers / detection scores
def _idx(t, name):
    if not t.is_cuda:
        raise RuntimeError("d3feat.pytorch_b200: `%s` must be a CUDA tensor" % name)
    if t.dtype not in (torch.int32, torch.int64):
        t = t.long()
    return t


def _valid_width(inds):
    """Device int32 tensor holding the reference's matrix width for a capacity-padded neighbour matrix
    (set by engine.collate_static), or None for exact-shape matrices."""
    return getattr(inds, "_d3f_width", None)


class _MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inds, width):
        lib = _lib.load()
        x, inds = _cuda_f32(x, "x"), _idx(inds, "inds")
        if inds.stride(-1) != 1:
            inds = inds.contiguous()
        nq, H = inds.shape
        ns, C = x.shape
        out = torch.empty((nq, C), dtype=torch.float32, device=x.device)
        arg = torch.empty((nq, C), dtype=torch.int32, device=x.device)
        global launch_count
        launch_count += 1
        with _Timed(("max_pool", nq, ns, H, C)):
            _lib.check(lib.d3f_max_pool_forward(_p(x), _p(inds), 1 if inds.dtype == torch.int64 else 0,
                                                inds.stride(0) if H > 0 else 0, nq, ns, H, C, _p(width), _p(out), _p(arg),
                                                _stream()))
        ctx.save_for_backward(arg)
        ctx.shape = (nq, ns, C)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (arg,) = ctx.saved_tensors
        nq, ns, C = ctx.shape
        g = _cuda_f32(g, "grad")
        gx = torch.empty((ns, C), dtype=torch.float32, device=g.device)
        with _Timed(("max_pool_bwd", nq, ns, C)):
            _lib.check(lib.d3f_max_pool_backward(_p(g), _p(arg), nq, ns, C, _p(gx), _stream()))
        return gx, None, None


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx):
        lib = _lib.load()
        x, idx = _cuda_f32(x, "x"), _idx(idx, "idx")
        if idx.dim() != 1:
            raise RuntimeError("gather_rows expects a 1-D index (a strided column view is fine)")
        m = idx.shape[0]
        ns, C = x.shape
        out = torch.empty((m, C), dtype=torch.float32, device=x.device)
        global launch_count
        launch_count += 1
        with _Timed(("gather_rows", m, ns, C)):
            _lib.check(lib.d3f_gather_rows_forward(_p(x), _p(idx), 1 if idx.dtype == torch.int64 else 0,
                                                   idx.stride(0) if m > 0 else 1, m, ns, C, _p(out), _stream()))
        ctx.save_for_backward(idx)
        ctx.shape = (m, ns, C)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (idx,) = ctx.saved_tensors
        m, ns, C = ctx.shape
        g = _cuda_f32(g, "grad")
        gx = torch.empty((ns, C), dtype=torch.float32, device=g.device)
        with _Timed(("gather_rows_bwd", m, ns, C)):
            _lib.check(lib.d3f_gather_rows_backward(_p(g), _p(idx), 1 if idx.dtype == torch.int64 else 0,
                                                    idx.stride(0) if m > 0 else 1, m, ns, C, _p(gx), _stream()))
        return gx, None


class _DetectionScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, neighbors, eval_mode, width):
        lib = _lib.load()
        feats, neighbors = _cuda_f32(feats, "features"), _idx(neighbors, "neighbors")
        if neighbors.stride(-1) != 1:
            neighbors = neighbors.contiguous()
        n, C = feats.shape
        H = neighbors.shape[1]
        scores = torch.empty((n, 1), dtype=torch.float32, device=feats.device)
        state = torch.empty(16, dtype=torch.uint8, device=feats.device)
        global launch_count
        launch_count += 1
        with _Timed(("det_scores", n, H, C)):
            _lib.check(lib.d3f_detection_scores_forward(_p(feats), _p(neighbors), 1 if neighbors.dtype == torch.int64 else 0,
                                                        neighbors.stride(0) if H > 0 else 0, n, H, C, int(eval_mode),
                                                        _p(width), _p(scores), _p(state), _stream()))
        ctx.save_for_backward(feats, neighbors, state, width)
        ctx.eval_mode = int(eval_mode)
        return scores

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        feats, neighbors, state, width = ctx.saved_tensors
        n, C = feats.shape
        H = neighbors.shape[1]
        g = _cuda_f32(g.reshape(-1), "grad")
        gf = torch.empty_like(feats)
        with _Timed(("det_scores_bwd", n, H, C)):
            _lib.check(lib.d3f_detection_scores_backward(_p(feats), _p(neighbors), 1 if neighbors.dtype == torch.int64 else 0,
                                                         neighbors.stride(0) if H > 0 else 0, n, H, C, ctx.eval_mode,
                                                         _p(width), _p(state), _p(g), _p(gf), _stream()))
        return gf, None, None, None


def max_pool(x, inds):
    return _MaxPool.apply(x, inds, _valid_width(inds))


def gather_rows(x, idx):
    return _GatherRows.apply(x, idx)


def detection_scores(feats, neighbors, eval_mode):
    return _DetectionScores.apply(feats, neighbors, eval_mode, _valid_width(neighbors))


# --------------------------------------------------------------------------- tensor-core GEMM (3xTF32) + fused linear
DST_SEEN = None   # optim.FlatSGD.verify_direct: set of id(param) whose gradient destination was handed to a kernel


def grad_dst(param):
    """Destination a backward kernel writes the gradient of `param` into (optim.FlatSGD(direct=True) attaches a view of
    its flat gradient buffer), or None: the gradient is then returned to autograd as usual."""
    dst = getattr(param, "_d3f_grad", None) if param is not None else None
    if dst is not None and DST_SEEN is not None:
        DST_SEEN.add(id(param))
    return dst


def _prezeroed(t):
    """True for a destination inside a gradient buffer that its owner keeps cleared between steps (FlatSGD zeroes the
    flat gradient in the optimiser kernel): kernels may accumulate into it without a zero fill of their own."""
    return t is not None and getattr(t, "_d3f_prezeroed", False)


class ZeroArena:
    """Outputs of the atomically combined (split-K) GEMMs of one step, carved from ONE buffer that is cleared by ONE fill
    at the top of the step instead of one memset node in front of every GEMM (58 of the 126 memset nodes of a captured pair
    step, each on the critical path of a latency-bound chain).  engine.PairStep installs it as ops.ZERO_ARENA around its
    body and calls reset() first; a view is handed out once per step and never reused inside it.  The buffer is sized by
    the previous step (capture() warms up eagerly first); a request that does not fit falls back to torch.zeros."""

    def __init__(self, device=None):
        self.device = device  # None: the current CUDA device at the first reset()
        self.buf = None
        self.off = 0          # floats handed out in this step
        self.spilled = 0      # floats that did not fit in this step
        self.high = 0         # floats handed out in the previous step (= what reset() has to clear)

    def reset(self):
        need = self.off + self.spilled
        if self.buf is None or self.buf.numel() < need:
            if self.device is None:
                self.device = torch.device("cuda", torch.cuda.current_device())
            self.buf = torch.zeros(need, dtype=torch.float32, device=self.device) if need else None
        elif self.off:
            self.buf[:self.off].zero_()
        self.high, self.off, self.spilled = self.off, 0, 0

    def take(self, n, device):
        n_al = (n + 31) & ~31                      # 128-byte granules: vector stores stay aligned
        if self.buf is None or self.buf.device != device or self.off + n_al > self.buf.numel():
            self.spilled += n_al
            return torch.zeros(n, dtype=torch.float32, device=device)
        v = self.buf[self.off:self.off + n]
        self.off += n_al
        return v


ZERO_ARENA = None


def gemm(a, b, trans_a=False, trans_b=False, row_scale=None, k_scale=None, bias=None, slope=None, deterministic=False,
         bias2=None, residual=None, out=None):
    """C = act(row_scale * opA(a) @ (k_scale * opB(b)) + bias + bias2 + residual) via d3f_gemm / d3f_gemm_ex
    (fp32-accurate tensor-core GEMM).  a: [M,K] (or [K,M] if trans_a); b: [K,N] (or [N,K] if trans_b).
    deterministic=True (forward pass): d3f_gemm_ex, whose split-K depends on K only and sums its partials in a fixed
    order; bias2 / residual need it."""
    lib = _lib.load()
    a, b = _cuda_f32(a, "a"), _cuda_f32(b, "b")
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    N = b.shape[0] if trans_b else b.shape[1]
    kb = b.shape[1] if trans_b else b.shape[0]
    if kb != K:
        raise RuntimeError("gemm: inner dimensions differ (%d vs %d)" % (K, kb))
    if out is not None:
        if out.numel() != M * N or out.dtype != torch.float32 or not out.is_contiguous():
            raise RuntimeError("gemm: `out` must be a contiguous fp32 tensor of %d elements" % (M * N))
        c = out
    elif (ZERO_ARENA is not None and not deterministic and bias is None and slope is None and bias2 is None
          and residual is None):
        # a plain GEMM may split K and combine its partials atomically: its output comes from the step's cleared arena
        c = out = ZERO_ARENA.take(M * N, a.device).view(M, N)
        c._d3f_prezeroed = True
    else:
        c = torch.empty((M, N), dtype=torch.float32, device=a.device)
    global launch_count
    launch_count += 1
    head = (int(trans_a), int(trans_b), M, N, K, _p(a), a.stride(0), _p(b), b.stride(0), _p(c), N,
            _p(None if row_scale is None else _cuda_f32(row_scale, "row_scale")),
            _p(None if k_scale is None else _cuda_f32(k_scale, "k_scale")),
            _p(None if bias is None else _cuda_f32(bias, "bias")))
    act = (0 if slope is None else 1, 0.0 if slope is None else float(slope))
    with _Timed(("gemm", M, N, K, bool(trans_a), bool(trans_b))):
        if deterministic or bias2 is not None or residual is not None:
            if residual is not None:
                residual = _cuda_f32(residual, "residual")
                if tuple(residual.shape) != (M, N):
                    raise RuntimeError("gemm: residual must be [%d, %d]" % (M, N))
            ws = _ws(lib.d3f_gemm_workspace_bytes(M, N, K), a.device)
            _lib.check(lib.d3f_gemm_ex(*head, _p(None if bias2 is None else _cuda_f32(bias2, "bias2")), _p(residual),
                                       N, *act, _p(ws), ws.numel(), _stream()))
        elif _prezeroed(out):
            _lib.check(lib.d3f_gemm_prezeroed(*head, *act, _stream()))
        else:
            _lib.check(lib.d3f_gemm(*head, *act, _stream()))
    return c


def colsum(x, out=None):
    """x.sum(dim=0) for a 2-D fp32 CUDA tensor (d3f_colsum)."""
    lib = _lib.load()
    x = _cuda_f32(x, "x")
    if out is None:
        out = torch.empty(x.shape[1], dtype=torch.float32, device=x.device)
    fn = lib.d3f_colsum_prezeroed if _prezeroed(out) else lib.d3f_colsum
    _lib.check(fn(_p(x), x.shape[0], x.shape[1], _p(out), _stream()))
    return out


# --------------------------------------------------------------------------- concurrent backward branches
# The weight-gradient GEMM of a layer (dW = dz^T x, or wf^T g for KPConv) and its data-gradient chain are independent
# and each is latency-bound on a few dozen CTAs: running the dW branch on an auxiliary stream lets them share the GPU.
# Inside a CUDA-graph capture the fork/join becomes two parallel branches of the graph.  BRANCHES = False runs them in
# order on the current stream.
BRANCHES = True
_AUX_STREAMS = {}


def _aux_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _AUX_STREAMS.get(key)
    if st is None:
        st = _AUX_STREAMS[key] = torch.cuda.Stream(device=key)
    return st


def run_branches(side_fn, main_fn, device):
    """side_fn() on the auxiliary stream concurrently with main_fn() on the current stream; both have completed (in
    stream order) when this returns.  Returns (side result, main result)."""
    if not BRANCHES:
        return side_fn(), main_fn()
    cur, aux = torch.cuda.current_stream(), _aux_stream(device)
    aux.wait_stream(cur)
    with torch.cuda.stream(aux):
        a = side_fn()
    b = main_fn()
    cur.wait_stream(aux)
    return a, b


def leaky_backward_colsum(gy, y, slope, want_colsum=True, db_out=None):
    """(dz, colsum(dz)) with dz = gy * (y > 0 ? 1 : slope), y = the saved LeakyReLU output: one kernel
    (d3f_leaky_backward_colsum) where the channel count allows it, else ATen's leaky_relu_backward + d3f_colsum."""
    lib = _lib.load()
    gy, y = _cuda_f32(gy, "grad"), _cuda_f32(y, "y")
    M, N = gy.shape
    if want_colsum:
        dz = torch.empty_like(gy)
        db = db_out if db_out is not None else torch.empty(N, dtype=torch.float32, device=gy.device)
        fn = lib.d3f_leaky_backward_colsum_prezeroed if _prezeroed(db_out) else lib.d3f_leaky_backward_colsum
        rc = fn(_p(gy), _p(y), float(slope), M, N, _p(dz), _p(db), _stream())
        if rc == 0:
            return dz, db
        if rc != -4:   # D3F_ERR_UNSUPPORTED: shape not covered by the vector kernel
            _lib.check(rc)
    dz = torch.ops.aten.leaky_relu_backward(gy, y, float(slope), True).contiguous()
    return dz, (colsum(dz, out=db_out) if want_colsum else None)


class _FusedLinear(torch.autograd.Function):
    """y = LeakyReLU_slope(x @ W^T + b + b2 + residual)  (slope None: no activation; b2 / residual optional) -- the
    UnaryBlock body (models/blocks.py:505-510 with use_bn=False: Linear bias + learned bias) and, with `residual`, the
    tail of ResnetBottleneckBlock (leaky_relu(unary2(x) + shortcut), blocks.py:686) as ONE GEMM with a fused epilogue."""

    @staticmethod
    def forward(ctx, x, weight, bias, bias2, residual, slope, dsts):
        y = gemm(x, weight, trans_b=True, bias=bias, slope=slope, deterministic=True, bias2=bias2, residual=residual)
        ctx.save_for_backward(x, weight, y if slope is not None else None)
        ctx.slope = slope
        ctx.dsts = dsts      # (dW, db, db2) destinations inside a flat gradient buffer, or Nones
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        need = ctx.needs_input_grad
        dst_w, dst_b, dst_b2 = ctx.dsts
        want_db = need[2] or need[3]
        db_out = dst_b if (need[2] and dst_b is not None) else (dst_b2 if (need[3] and dst_b2 is not None) else None)
        if y is None:
            dz = gy.contiguous()
            db = colsum(dz, out=db_out) if want_db else None
        else:   # y = leaky(z) has the sign of z (slope > 0): the mask comes from the saved output, fused with the bias grad
            dz, db = leaky_backward_colsum(gy, y, ctx.slope, want_db, db_out)
        if need[0] and need[1]:
            dw, dx = run_branches(lambda: gemm(dz, x, trans_a=True, out=dst_w),    # dz^T [out,M] @ x [M,in]
                                  lambda: gemm(dz, weight), dz.device)             # [M,out] @ [out,in]
        else:
            dx = gemm(dz, weight) if need[0] else None
            dw = gemm(dz, x, trans_a=True, out=dst_w) if need[1] else None
        gb = gb2 = None
        if want_db:
            if db_out is None:                      # plain autograd gradients
                gb, gb2 = (db if need[2] else None), (db if need[3] else None)
            elif db_out is dst_b:                   # written in place; the second bias gets the same column sums
                if need[3]:
                    if dst_b2 is not None:
                        dst_b2.copy_(db)
                    else:
                        gb2 = db.clone()
            else:                                   # db_out is dst_b2
                if need[2]:
                    gb = db.clone()
        if dst_w is not None:
            dw = None
        return dx, dw, gb, gb2, dz if need[4] else None, None, None


def fused_linear(x, weight, bias, slope=None, bias2=None, residual=None):
    return _FusedLinear.apply(x, weight, bias, bias2, residual, slope, (grad_dst(weight), grad_dst(bias), grad_dst(bias2)))
