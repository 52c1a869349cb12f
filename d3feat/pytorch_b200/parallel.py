"""Multi-GPU plumbing: one process per GPU, one fragment pair per rank.

The hot path shards by pair with no data-path collective except ONE exchange step: a packed
all-gather of every rank's P selected descriptors / scores / keypoint distances, so that each rank
can evaluate the cross-fragment (B*P x B*P) descriptor + detector loss.  The reference has no
multi-GPU code; semantics follow SURVEY.md 8(e): positives = global diagonal, entries between
different pairs are always valid negatives (dist_keypts = +inf off the block diagonal), and
B = 1 reproduces the single-pair loss exactly.

``FlatGradients`` keeps all parameter gradients as views of one flat buffer so the data-parallel
gradient exchange is a single in-place all-reduce (SUM: every rank differentiates the SAME global
loss with respect to its own slice of the gathered descriptors, so the per-rank weight gradients
add up to the gradient of that loss).
"""
import torch
import torch.distributed as dist


def _pack(a, p, sa, sp, dk):
    """[P,D]+[P,D]+[P]+[P] f32 and [P,P] f64 -> one uint8 buffer (one collective instead of five)."""
    f32 = torch.cat([a.detach().reshape(-1), p.detach().reshape(-1), sa.detach().reshape(-1), sp.detach().reshape(-1)]).float()
    return torch.cat([f32.contiguous().view(torch.uint8), dk.detach().double().contiguous().view(-1).view(torch.uint8)])


def _unpack(buf, P, D):
    n32 = (2 * P * D + 2 * P) * 4
    f32 = buf[:n32].view(torch.float32)
    a, p = f32[:P * D].view(P, D), f32[P * D:2 * P * D].view(P, D)
    sa, sp = f32[2 * P * D:2 * P * D + P].view(P, 1), f32[2 * P * D + P:].view(P, 1)
    dk = buf[n32:].view(torch.float64).view(P, P)
    return a, p, sa, sp, dk


class _AttachLocal(torch.autograd.Function):
    """`gathered` [W*P, ...] holds every rank's rows (constants); rows lo:lo+P are bit-identical copies of `local`.
    Returns `gathered` attached to the autograd graph of `local`: the gradient of the local slice flows back, the other
    ranks' rows stay constants (what torch.cat of [const..., local, const...] did, without the W+1 kernels)."""

    @staticmethod
    def forward(ctx, local, gathered, lo):
        ctx.lo, ctx.shape = lo, local.shape
        ctx.mark_dirty(gathered)
        return gathered

    @staticmethod
    def backward(ctx, g):
        n = ctx.shape[0]
        return g[ctx.lo:ctx.lo + n].reshape(ctx.shape), None, None


def _gather_pairs_cuda(a, p, sa, sp, dk, group):
    """gather_pairs on the GPU: ONE pack kernel, ONE all-gather, ONE unpack kernel (csrc/exchange.cu)."""
    from . import _lib
    lib = _lib.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    P, D = a.shape
    dev = a.device
    st = torch.cuda.current_stream().cuda_stream
    chunk = int(lib.d3f_exchange_chunk_bytes(P, D))
    mine = torch.empty(chunk, dtype=torch.uint8, device=dev)
    af, pf = a.detach().float().contiguous(), p.detach().float().contiguous()
    saf, spf = sa.detach().float().reshape(-1).contiguous(), sp.detach().float().reshape(-1).contiguous()
    if dk.dtype not in (torch.float32, torch.float64):
        dk = dk.double()
    dkc = dk.detach().contiguous()
    _lib.check(lib.d3f_exchange_pack(af.data_ptr(), pf.data_ptr(), saf.data_ptr(), spf.data_ptr(), dkc.data_ptr(),
                                     int(dkc.dtype == torch.float64), P, D, mine.data_ptr(), st))
    allbuf = torch.empty(world * chunk, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allbuf, mine, group=group)
    A = torch.empty((world * P, D), dtype=torch.float32, device=dev)
    Pos = torch.empty_like(A)
    SA = torch.empty((world * P, 1), dtype=torch.float32, device=dev)
    SP = torch.empty_like(SA)
    DK = torch.empty((world * P, world * P), dtype=torch.float64, device=dev)
    _lib.check(lib.d3f_exchange_unpack(allbuf.data_ptr(), world, P, D, A.data_ptr(), Pos.data_ptr(), SA.data_ptr(),
                                       SP.data_ptr(), DK.data_ptr(), st))
    lo = rank * P
    attach = lambda local, full: _AttachLocal.apply(local, full, lo) if local.requires_grad else full
    return (attach(a, A), attach(p, Pos), attach(sa.reshape(P, 1), SA), attach(sp.reshape(P, 1), SP), DK)


def gather_pairs(a, p, sa, sp, dk, group=None):
    """All-gather every rank's (anchor, positive, scores, dist_keypts).  The local rank's slices stay
    attached to the autograd graph; the other ranks' slices are constants.
    Returns (A [W*P,D], Pos [W*P,D], SA [W*P,1], SP [W*P,1], DK [W*P,W*P] f64 block-diagonal, +inf elsewhere).
    CUDA tensors take the two-kernel path of csrc/exchange.cu; CPU tensors (the gloo tests of the host logic) the
    torch-op restatement below."""
    if a.is_cuda and a.dtype == torch.float32 and p.dtype == torch.float32:
        return _gather_pairs_cuda(a, p, sa, sp, dk, group)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    P, D = a.shape
    mine = _pack(a, p, sa, sp, dk)
    allbuf = torch.empty(world * mine.numel(), dtype=torch.uint8, device=mine.device)
    dist.all_gather_into_tensor(allbuf, mine, group=group)
    As, Ps, SAs, SPs = [], [], [], []
    DK = torch.full((world * P, world * P), float("inf"), dtype=torch.float64, device=a.device)
    for r in range(world):
        ra, rp, rsa, rsp, rdk = _unpack(allbuf[r * mine.numel():(r + 1) * mine.numel()], P, D)
        if r == rank:
            ra, rp, rsa, rsp = a, p, sa.reshape(P, 1), sp.reshape(P, 1)
        As.append(ra); Ps.append(rp); SAs.append(rsa); SPs.append(rsp)
        DK[r * P:(r + 1) * P, r * P:(r + 1) * P] = rdk
    return torch.cat(As), torch.cat(Ps), torch.cat(SAs), torch.cat(SPs), DK


def cross_fragment_loss(loss_fn, a, p, dist_keypts, sa, sp, group=None):
    """loss_fn(A, Pos, DK, SA, SP) on the gathered batch (PairLoss on the GPU)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return loss_fn(a, p, dist_keypts, sa, sp)
    A, Pos, SA, SP, DK = gather_pairs(a, p, sa, sp, dist_keypts, group)
    return loss_fn(A, Pos, DK, SA, SP)


class FlatGradients:
    """Parameter gradients as views of one contiguous buffer; `zero()` replaces optimizer.zero_grad()
    and `allreduce()` is the whole data-parallel gradient exchange."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()

    def allreduce(self, group=None):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)


def allreduce_gradients(module, group=None):
    """One-shot variant for modules without a FlatGradients: flatten, all-reduce (SUM), copy back."""
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()
