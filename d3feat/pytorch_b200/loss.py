"""Descriptor / detector losses -- drop-in for utils/loss.py of the reference, computed by the
sm_100a pair-loss kernels (d3f_pair_loss_forward / _backward).

``CircleLoss``, ``ContrastiveLoss`` and ``DetLoss`` keep the reference signatures and return
values (loss.py:111-141, 55-97, 149-158).  ``PairLoss`` is the fused form of the trainer wiring
(trainer.py:96-98): descriptor loss + detector loss in one forward/backward, no host sync.
"""
import torch
import torch.nn as nn

from . import _lib, ops


def cdist(a, b, metric='euclidean'):
    """Pairwise distance matrix with the reference's metrics (loss.py:8-44).  Forward only."""
    if metric not in ops.METRIC:
        raise NotImplementedError('The following metric is not implemented by `cdist` yet: {}'.format(metric))
    return ops.pair_dist(a, b, metric)


class _PairLossFunction(torch.autograd.Function):
    """(anchor, positive, anc_score, pos_score) -> (desc_loss, det_loss, dists, accuracy, furthest_pos, avg_neg)."""

    @staticmethod
    def forward(ctx, anchor, positive, anc_score, pos_score, dist_keypts, kind, metric, safe_radius, pos_margin,
                neg_margin, log_scale):
        dists, stats, fp, an, aux, saved = ops.pair_loss_forward(anchor, positive, dist_keypts, anc_score, pos_score,
                                                                 kind, metric, safe_radius, pos_margin, neg_margin,
                                                                 log_scale)
        ctx.set_materialize_grads(False)
        anchor_c, positive_c, dk, sa, sp = saved
        # `dists` is an output: it must go through save_for_backward (a plain ctx attribute would form a
        # reference cycle ctx -> dists -> grad_fn that keeps the whole autograd graph alive until the GC runs)
        ctx.save_for_backward(anchor_c, positive_c, dk, sa, sp, dists, aux)
        ctx.cfg = (kind, metric, safe_radius, pos_margin, neg_margin, log_scale)
        ctx.score_shapes = (None if anc_score is None else anc_score.shape,
                            None if pos_score is None else pos_score.shape)
        acc = stats[2].clone()
        ctx.mark_non_differentiable(acc, fp, an)
        return stats[0].clone(), stats[1].clone(), dists, acc, fp, an

    @staticmethod
    def backward(ctx, g_desc, g_det, g_dists, _a, _b, _c):
        if g_dists is not None:
            raise RuntimeError("the `dists` output of the fused pair loss only feeds DetLoss inside the kernel; "
                               "use PairLoss (or DetLoss on it) rather than differentiating it directly")
        anchor_c, positive_c, dk, sa, sp, dists, aux = ctx.saved_tensors
        gl = torch.stack([g_desc if g_desc is not None else torch.zeros((), device=aux.device),
                          g_det if g_det is not None else torch.zeros((), device=aux.device)]).float()
        kind, metric, safe_radius, pm, nm, ls = ctx.cfg
        ga, gp, gsa, gsp = ops.pair_loss_backward((anchor_c, positive_c, dk, sa, sp), kind, metric, safe_radius, pm, nm, ls,
                                                  dists, aux, gl)
        sa_shape, sp_shape = ctx.score_shapes
        gsa = gsa.reshape(sa_shape) if gsa is not None else None
        gsp = gsp.reshape(sp_shape) if gsp is not None else None
        return ga, gp, gsa, gsp, None, None, None, None, None, None, None


class PairLoss(nn.Module):
    """Descriptor loss ('circle' | 'contrastive') + detector loss on its distance matrix, fused.

    forward(anchor [P,D], positive [P,D], dist_keypts [P,P], anc_score [P,1], pos_score [P,1])
      -> dict(desc_loss, det_loss, accuracy, dists, furthest_positive, average_negative)   (device tensors)
    """

    def __init__(self, desc_loss='circle', dist_type='euclidean', log_scale=10, safe_radius=0.10, pos_margin=0.1,
                 neg_margin=1.4):
        super().__init__()
        self.kind, self.metric = desc_loss, dist_type
        self.log_scale, self.safe_radius = log_scale, safe_radius
        self.pos_margin, self.neg_margin = pos_margin, neg_margin

    def forward(self, anchor, positive, dist_keypts, anc_score=None, pos_score=None):
        desc, det, dists, acc, fp, an = _PairLossFunction.apply(
            anchor, positive, anc_score, pos_score, dist_keypts, self.kind, self.metric, self.safe_radius,
            self.pos_margin, self.neg_margin, self.log_scale)
        return dict(desc_loss=desc, det_loss=det, accuracy=acc, dists=dists, furthest_positive=fp,
                    average_negative=an)


def _tag(dists, **meta):
    """Remember the inputs `dists` was computed from, so that DetLoss(dists, ...) can run the fused
    kernel: detector gradients then reach the descriptors, as they do through the reference's
    un-detached `dists` (loss.py:141 / trainer.py:96-97)."""
    dists = dists.detach()
    dists._d3f_meta = meta
    return dists


class CircleLoss(nn.Module):
    def __init__(self, dist_type='cosine', log_scale=10, safe_radius=0.10, pos_margin=0.1, neg_margin=1.4):
        super().__init__()
        self.log_scale = log_scale
        self.pos_margin = pos_margin
        self.neg_margin = neg_margin
        self.pos_optimal = pos_margin
        self.neg_optimal = neg_margin
        self.dist_type = dist_type
        self.safe_radius = safe_radius
        self._kind = 'circle'

    def forward(self, anchor, positive, dist_keypts):
        desc, _det, dists, acc, fp, an = _PairLossFunction.apply(
            anchor, positive, None, None, dist_keypts, self._kind, self.dist_type, self.safe_radius,
            self.pos_margin, self.neg_margin, self.log_scale)
        dists = _tag(dists, anchor=anchor, positive=positive, dist_keypts=dist_keypts, kind=self._kind,
                     metric=self.dist_type, safe_radius=self.safe_radius, pos_margin=self.pos_margin,
                     neg_margin=self.neg_margin, log_scale=self.log_scale)
        return desc, acc, fp.tolist(), an.tolist(), 0, dists


class ContrastiveLoss(nn.Module):
    def __init__(self, pos_margin=0.1, neg_margin=1.4, metric='euclidean', safe_radius=0.25):
        super().__init__()
        self.pos_margin = pos_margin
        self.neg_margin = neg_margin
        self.metric = metric
        self.safe_radius = safe_radius

    def forward(self, anchor, positive, dist_keypts):
        desc, _det, dists, acc, fp, an = _PairLossFunction.apply(
            anchor, positive, None, None, dist_keypts, 'contrastive', self.metric, self.safe_radius,
            self.pos_margin, self.neg_margin, 10.0)
        dists = _tag(dists, anchor=anchor, positive=positive, dist_keypts=dist_keypts, kind='contrastive',
                     metric=self.metric, safe_radius=self.safe_radius, pos_margin=self.pos_margin,
                     neg_margin=self.neg_margin, log_scale=10.0)
        return desc, acc, fp.tolist(), an.tolist(), 0, dists


class _DetLossFunction(torch.autograd.Function):
    """utils/loss.py:149-158 on an arbitrary distance matrix (d3f_det_loss_forward / _backward)."""

    @staticmethod
    def forward(ctx, dists, anc_score, pos_score):
        lib = _lib.load()
        if not dists.is_cuda:
            raise RuntimeError("DetLoss: CUDA tensors only (there is no CPU path)")
        d = dists.detach().float().contiguous()
        P = d.shape[0]
        if d.dim() != 2 or d.shape[1] != P or anc_score.numel() != P or pos_score.numel() != P:
            raise RuntimeError("DetLoss: dists must be [P,P] and the scores [P] or [P,1]")
        a = anc_score.detach().float().reshape(-1).contiguous()
        b = pos_score.detach().float().reshape(-1).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=d.device)
        rowval = torch.empty(P, dtype=torch.float32, device=d.device)
        arg = torch.empty(2 * P, dtype=torch.int32, device=d.device)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.d3f_det_loss_forward(d.data_ptr(), P, a.data_ptr(), b.data_ptr(), P, loss.data_ptr(),
                                            rowval.data_ptr(), arg.data_ptr(), st))
        ctx.save_for_backward(rowval, arg, a, b)
        ctx.shapes = (dists.shape, anc_score.shape, pos_score.shape, dists.dtype)
        return loss[0]

    @staticmethod
    def backward(ctx, grad):
        lib = _lib.load()
        rowval, arg, a, b = ctx.saved_tensors
        P = rowval.shape[0]
        dshape, ashape, bshape, ddtype = ctx.shapes
        g = grad.detach().float().reshape(1).contiguous()
        gd = torch.empty((P, P), dtype=torch.float32, device=g.device) if ctx.needs_input_grad[0] else None
        ga = torch.empty(P, dtype=torch.float32, device=g.device) if ctx.needs_input_grad[1] else None
        gb = torch.empty(P, dtype=torch.float32, device=g.device) if ctx.needs_input_grad[2] else None
        _lib.check(lib.d3f_det_loss_backward(rowval.data_ptr(), arg.data_ptr(), a.data_ptr(), b.data_ptr(), P, g.data_ptr(),
                                             None if gd is None else gd.data_ptr(), P,
                                             None if ga is None else ga.data_ptr(), None if gb is None else gb.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream))
        return (None if gd is None else gd.to(ddtype).reshape(dshape), None if ga is None else ga.reshape(ashape),
                None if gb is None else gb.reshape(bshape))


class DetLoss(nn.Module):
    """DetLoss.forward(dists, anc_score, pos_score) of utils/loss.py:144-158.

    `dists` returned by this package's CircleLoss / ContrastiveLoss carries the inputs it was computed from: the fused
    pair-loss kernels then run again with the scores, so detector gradients reach the descriptors as they do through the
    reference's un-detached `dists` (trainer.py:96-97).  Any OTHER [P,P] CUDA matrix (a clone, a slice, a cdist) takes
    the stand-alone detector-loss kernels, differentiable in `dists` and both scores, as in the reference."""

    def __init__(self, metric='euclidean'):
        super().__init__()
        self.metric = metric

    def forward(self, dists, anc_score, pos_score):
        meta = getattr(dists, '_d3f_meta', None)
        if meta is None:
            return _DetLossFunction.apply(dists, anc_score, pos_score)
        _desc, det, *_ = _PairLossFunction.apply(
            meta['anchor'], meta['positive'], anc_score, pos_score, meta['dist_keypts'], meta['kind'],
            meta['metric'], meta['safe_radius'], meta['pos_margin'], meta['neg_margin'], meta['log_scale'])
        return det
