"""Optimiser step of the reference trainer on flat device buffers, with no host round trip:

* SGD with momentum 0.98 / weight decay 1e-6 / lr 0.01 (training_3DMatch.py:62-69, config.py:67-69),
* ExponentialLR, gamma = 0.1 ** (1/80), stepped once per epoch (training_3DMatch.py:77-80, trainer.py:60-61):
  the learning rate lives in device memory and `scheduler_step()` multiplies it in place, so a captured CUDA graph
  sees the new value without being re-captured,
* the non-finite-gradient guard of trainer.py:104-111 ("skip optimizer.step() if any gradient has an inf / nan") as a
  device-side predicate: `d3f_sgd_step` scans the flat gradient, raises a flag and the update kernel returns early.

All parameters (and gradients, and momentum buffers) are views of ONE flat buffer each, ordered so that the tensors whose
gradients are finished first by the backward pass (decoder + the deep encoder levels, 90 % of the weights) form a
contiguous prefix: the data-parallel all-reduce of that bucket is launched on a side stream while the shallow levels are
still back-propagating (engine.PairStep), the rest follows at the end.

Gradients are WRITTEN into their flat slice by the kernels that compute them (every parameter of KPFCNN is produced by
one of this package's autograd Functions, which take the destination through `param._d3f_grad`): no zero-fill, no
per-parameter accumulate kernel.  `verify_direct()` proves that property for a model (it records which destinations the
kernels of one step actually fetched).
"""
import torch
import torch.distributed as dist

from . import _lib


def _p(t):
    return None if t is None else t.data_ptr()


class FlatSGD:
    def __init__(self, module, lr=0.01, momentum=0.98, weight_decay=1e-6, gamma=0.1 ** (1 / 80), early=None,
                 direct=True):
        """`early(name) -> bool` selects the parameters of the first all-reduce bucket (default: everything except the
        encoder blocks of the three shallowest levels, see `default_early`)."""
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if not named:
            raise ValueError("FlatSGD: the module has no trainable parameters")
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("d3feat.pytorch_b200: FlatSGD needs CUDA parameters (there is no CPU path)")
        early = early if early is not None else default_early(module)
        first = [(n, p) for n, p in named if early(n)]
        second = [(n, p) for n, p in named if not early(n)]
        self.names = [n for n, _ in first + second]
        self.params = [p for _, p in first + second]
        # 16-byte aligned slices (vector kernels write straight into them)
        offs, o = [], 0
        for p in self.params:
            offs.append(o)
            o += (p.numel() + 3) // 4 * 4
        self.n = o
        self.split = offs[len(first)] if second else o
        self.flat_p = torch.zeros(o, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(o, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(o, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, offs):
            view = self.flat_p[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            g = self.flat_g[off:off + p.numel()].view_as(p)
            p.grad = g
            if direct:
                g._d3f_prezeroed = True  # zero_grad() clears the flat gradient once per step: kernels may accumulate into it
                p._d3f_grad = g          # destination the backward kernels write into (ops.grad_dst)
        self.direct = direct
        self.lr = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.base_lr, self.gamma, self.epoch = float(lr), float(gamma), 0
        self.momentum, self.weight_decay = float(momentum), float(weight_decay)
        self.nonfinite = torch.zeros(1, dtype=torch.int32, device=dev)   # 1 = the last step was skipped
        self.comm_stream = None
        self._early_pending = False

    # -- torch.optim-like surface used by engine.PairStep
    def zero_grad(self, set_to_none=False):
        """ONE fill of the flat gradient per step.  With direct=True that is what lets the backward kernels accumulate
        split-K / row-block partial sums into their slices without ~100 per-kernel zero fills; engine.PairStep issues it
        at the top of the step, where the main stream is idle waiting for the first radius search.  (Clearing the
        gradient inside the optimiser kernel instead was measured 4.5x slower: 377 vs 84 us for 24.3 M parameters.)"""
        self.flat_g.zero_()

    def step(self):
        lib = _lib.load()
        self.nonfinite.zero_()
        _lib.check(lib.d3f_sgd_step(_p(self.flat_p), _p(self.flat_g), _p(self.flat_m), self.n, _p(self.lr),
                                    self.momentum, self.weight_decay, _p(self.nonfinite), 1, 0,
                                    torch.cuda.current_stream().cuda_stream))

    def scheduler_step(self):
        """ExponentialLR.step(): lr <- lr * gamma (in device memory; valid for an already captured graph)."""
        self.epoch += 1
        self.lr.mul_(self.gamma)

    def current_lr(self):
        return self.base_lr * self.gamma ** self.epoch

    # -- data-parallel gradient exchange (SUM, see parallel.py), in two buckets
    def allreduce_early(self, group=None):
        """Called from a backward hook once the gradients of the first bucket are complete (engine.PairStep)."""
        if not (dist.is_initialized() and dist.get_world_size(group) > 1) or self.split == 0:
            return
        cur = torch.cuda.current_stream()
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=self.flat_g.device)
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(self.flat_g[:self.split], op=dist.ReduceOp.SUM, group=group)
        self._early_pending = True

    def allreduce(self, group=None):
        """The rest of the exchange, after backward: second bucket (or everything if the early bucket was not sent)."""
        if not (dist.is_initialized() and dist.get_world_size(group) > 1):
            return
        cur = torch.cuda.current_stream()
        if self._early_pending:
            if self.split < self.n:
                dist.all_reduce(self.flat_g[self.split:], op=dist.ReduceOp.SUM, group=group)
            cur.wait_stream(self.comm_stream)
            self._early_pending = False
        else:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=group)

    def verify_direct(self, run_step):
        """Prove that one step hands EVERY parameter's gradient slice to the kernel that computes it (so autograd never
        accumulates into the flat buffer behind our back and nothing needs a per-parameter zero fill): record the
        destinations fetched during `run_step()` (ops.grad_dst) and compare with the parameter list."""
        from . import ops
        ops.DST_SEEN = set()
        try:
            run_step()
            torch.cuda.synchronize()
            seen = ops.DST_SEEN
        finally:
            ops.DST_SEEN = None
        missing = [n for n, p in zip(self.names, self.params) if id(p) not in seen]
        if missing:
            raise RuntimeError("FlatSGD(direct=True): no kernel took the gradient destination of %s" % missing[:8])


EARLY_LEVEL = 3   # encoder blocks of pyramid levels >= 3 (and the whole decoder) form the first all-reduce bucket


def early_block_index(module):
    """Index of the first encoder block of level EARLY_LEVEL: when the gradient of ITS INPUT is ready, every Function
    of the decoder and of the deeper encoder blocks has run its backward (encoder and decoder are chains), i.e. the
    first bucket is complete.  None if the network has no such block."""
    for i, b in enumerate(getattr(module, "encoder_blocks", [])):
        if getattr(b, "layer_ind", -1) >= EARLY_LEVEL:
            return i
    return None


def default_early(module):
    """First bucket = decoder + encoder blocks from `early_block_index` on (their backward finishes first and they
    hold ~90 % of the weights: architectures.py:213-294, widths double per level)."""
    first = early_block_index(module)
    n_enc = len(getattr(module, "encoder_blocks", []))
    shallow = tuple("encoder_blocks.%d." % i for i in range(first if first is not None else n_enc))

    def early(name):
        return first is not None and not name.startswith(shallow)
    return early
