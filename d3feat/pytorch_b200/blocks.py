"""nn.Module surface of the KPConv hot path -- drop-in for models/blocks.py of the reference.

Same class names, constructor signatures, attribute names and state_dict keys as the
reference; the body of ``KPConv.forward`` is a torch.autograd.Function
over the sm_100a kernels behind the C ABI (d3f_kpconv_forward / d3f_kpconv_backward).
The non-hot-path blocks (unary MLPs, pooling, upsampling) stay stock PyTorch, as in the
reference.  There is no CPU path: KPConv raises on CPU tensors.
"""
import math

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import ops
from .kernel_points import load_kernels


# ----------------------------------------------------------------------------- small tensor helpers
def gather(x, idx, method=2):
    """x[idx] (reference blocks.py:35-66 offers three equivalent formulations; one is enough here)."""
    if idx.dim() == 1 and x.dim() == 2 and x.is_cuda:
        return ops.gather_rows(x, idx)
    return x[idx.long()]


def closest_pool(x, inds):
    """Feature of the closest (first-column) neighbour; shadow -> zeros (blocks.py:79-91).
    One row-gather kernel; the backward is an atomic row scatter instead of ATen's sort-based index_put."""
    return ops.gather_rows(x, inds[:, 0])


def max_pool(x, inds):
    """Channel-wise max over each pooling neighbourhood; the shadow row is zero (blocks.py:94-110).
    One warp-per-query kernel that keeps the arg-max rows for the backward scatter."""
    return ops.max_pool(x, inds)


def global_average(x, batch_lengths):
    """Per-cloud mean of stacked features (blocks.py:113-133)."""
    lengths = [int(v) for v in batch_lengths]
    return torch.stack([chunk.mean(dim=0) for chunk in torch.split(x, lengths)])


# ----------------------------------------------------------------------------- KPConv
class _KPConvFunction(torch.autograd.Function):
    """out = act(KPConv(q, s, inds, x; W, kernel points [, modulations]) + bias) through libd3feat_b200
    (bias / slope None: the bare convolution)."""

    @staticmethod
    def forward(ctx, q_pts, s_pts, inds, x, weights, kpoints, modulations, extent, influence, aggregation,
                deformed, want_min_d2, bias=None, slope=None, t_off=None, t_src=None, dsts=(None, None)):
        # with transposed neighbour lists a rigid layer's backward works from the gather over those lists alone (grad_x AND
        # grad_W): the kernel-point-weighted features wf are then not needed, and the fused kernel never writes them
        lists = (t_off is not None and t_src is not None and inds.shape[1] > 0 and q_pts.shape[0] > 0 and s_pts.shape[0] > 0
                 and ops.kpconv_fused_eligible(inds.shape[1], weights.shape[0], weights.shape[1], weights.shape[2],
                                               deformed, modulations, influence, aggregation))
        out, wf, wf_un, inv_n, min_d2 = ops.kpconv_forward(q_pts, s_pts, inds, x, weights, kpoints, extent,
                                                           influence, aggregation, deformed, modulations,
                                                           want_min_d2, bias, slope, need_wf=not lists)
        if lists:
            wf = None
        ctx.save_for_backward(q_pts, s_pts, inds, x, weights, kpoints, modulations, wf, wf_un, inv_n,
                              out if slope is not None else None)
        ctx.cfg = (extent, influence, aggregation, deformed, slope)
        ctx.transpose = (t_off, t_src) if (t_off is not None and t_src is not None) else None
        ctx.lists = lists
        ctx.dsts = dsts     # (grad_weights, grad_bias) destinations inside a flat gradient buffer, or Nones
        if min_d2 is None:
            min_d2 = out.new_empty(0)
        ctx.mark_non_differentiable(min_d2)
        return out, min_d2

    @staticmethod
    def backward(ctx, grad_out, _grad_min_d2):
        q_pts, s_pts, inds, x, weights, kpoints, modulations, wf, wf_un, inv_n, out = ctx.saved_tensors
        extent, influence, aggregation, deformed, slope = ctx.cfg
        need = ctx.needs_input_grad
        want_gb = len(need) > 12 and need[12]
        dst_w, dst_b = ctx.dsts
        if slope is not None:   # out = leaky(z) has the sign of z: the mask comes from the saved output
            grad_out, gbias = ops.leaky_backward_colsum(grad_out, out, slope, want_gb, dst_b if want_gb else None)
        else:
            gbias = ops.colsum(grad_out, out=dst_b) if want_gb else None
        if dst_b is not None:
            gbias = None            # written in place
        args = (q_pts.float().contiguous(), s_pts.float().contiguous(),
                inds if inds.dtype in (torch.int32, torch.int64) else inds.long(),
                x.float().contiguous(), weights.contiguous(), kpoints.float().contiguous(), extent, influence,
                aggregation, deformed, modulations, wf, wf_un, inv_n, grad_out.contiguous())
        need_data = need[3] or (need[5] and deformed) or need[6]
        if ctx.lists:
            # fused-forward layers kept no wf: G (the gather over the transposed lists) feeds both gradients, and the two
            # GEMMs that consume it run as concurrent branches
            G = ops.kpconv_gather_transposed(args[0], args[1], ctx.transpose, args[14], inv_n, args[5], weights.shape[2],
                                             extent, influence, aggregation)
            gkp = gmod = None
            if need[3] and need[4]:
                (_, gw), (gx, _) = ops.run_branches(
                    lambda: ops.kpconv_grads_from_gathered(G, args[3], args[4], False, True, dst_w),
                    lambda: ops.kpconv_grads_from_gathered(G, args[3], args[4], True, False), grad_out.device)
            else:
                gx, gw = ops.kpconv_grads_from_gathered(G, args[3], args[4], need[3], need[4], dst_w)
        elif need[4] and need_data:
            # the weight gradient (one GEMM over wf) and the data-gradient chain are independent: two branches
            (_, gw, _, _), (gx, _, gkp, gmod) = ops.run_branches(
                lambda: ops.kpconv_backward(*args, need_x=False, need_w=True, need_kp=False, need_mod=False, gw_out=dst_w),
                lambda: ops.kpconv_backward(*args, need_x=need[3], need_w=False, need_kp=need[5] and deformed,
                                            need_mod=need[6], transpose=ctx.transpose), grad_out.device)
        else:
            gx, gw, gkp, gmod = ops.kpconv_backward(*args, need_x=need[3], need_w=need[4], need_kp=need[5] and deformed,
                                                    need_mod=need[6], transpose=ctx.transpose, gw_out=dst_w)
        if dst_w is not None:
            gw = None               # written in place
        return (None, None, None, gx, gw, gkp, gmod, None, None, None, None, None, gbias, None, None, None, None)[:len(need)]


class KPConv(nn.Module):
    """Kernel point convolution, rigid or deformable (reference: models/blocks.py:143-387)."""

    def __init__(self, kernel_size, p_dim, in_channels, out_channels, KP_extent, radius,
                 fixed_kernel_points='center', KP_influence='linear', aggregation_mode='sum',
                 deformable=False, modulated=False):
        super().__init__()
        if KP_influence not in ops.INFLUENCE:
            raise ValueError('Unknown influence function type (config.KP_influence)')
        if aggregation_mode not in ops.AGGREGATION:
            raise ValueError("Unknown convolution mode. Should be 'closest' or 'sum'")
        self.K = kernel_size
        self.p_dim = p_dim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.radius = radius
        self.KP_extent = KP_extent
        self.fixed_kernel_points = fixed_kernel_points
        self.KP_influence = KP_influence
        self.aggregation_mode = aggregation_mode
        self.deformable = deformable
        self.modulated = modulated

        # side outputs of the deformable path (blocks.py:177-180)
        self.min_d2 = None
        self.deformed_KP = None
        self.offset_features = None

        self.weights = Parameter(torch.zeros((self.K, in_channels, out_channels), dtype=torch.float32))
        if deformable:
            self.offset_dim = (self.p_dim + 1 if modulated else self.p_dim) * self.K
            self.offset_conv = KPConv(self.K, self.p_dim, self.in_channels, self.offset_dim, KP_extent, radius,
                                      fixed_kernel_points=fixed_kernel_points, KP_influence=KP_influence,
                                      aggregation_mode=aggregation_mode)
            self.offset_bias = Parameter(torch.zeros(self.offset_dim, dtype=torch.float32))
        else:
            self.offset_dim = None
            self.offset_conv = None
            self.offset_bias = None
        self.reset_parameters()
        self.kernel_points = self.init_KP()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))
        if self.deformable:
            nn.init.zeros_(self.offset_bias)

    def init_KP(self):
        kp = load_kernels(self.radius, self.K, dimension=self.p_dim, fixed=self.fixed_kernel_points)
        return Parameter(torch.tensor(kp, dtype=torch.float32), requires_grad=False)

    def forward(self, q_pts, s_pts, neighb_inds, x, bias=None, slope=None):
        """Reference signature forward(q_pts, s_pts, neighb_inds, x).  The optional `bias` [out_channels] and `slope`
        fuse the `+ bias` / LeakyReLU(slope) that the calling block applies next into the contraction's epilogue."""
        kpoints, modulations, deformed = self.kernel_points, None, False
        if self.deformable:
            # offsets come from a rigid KPConv over the same neighbourhoods (blocks.py:243-266)
            self.offset_features = self.offset_conv(q_pts, s_pts, neighb_inds, x, bias=self.offset_bias)
            n_off = self.p_dim * self.K
            unscaled = self.offset_features[:, :n_off].reshape(-1, self.K, self.p_dim)
            if self.modulated:
                modulations = 2 * torch.sigmoid(self.offset_features[:, n_off:])
            self.deformed_KP = unscaled * self.KP_extent + self.kernel_points
            kpoints, deformed = self.deformed_KP, True
        # transposed neighbour lists attached by engine.collate_static (ops.neighbors_transpose): atomic-free backward
        t_off, t_src = getattr(neighb_inds, "_d3f_transpose", (None, None))
        out, min_d2 = _KPConvFunction.apply(q_pts, s_pts, neighb_inds, x, self.weights, kpoints, modulations,
                                            float(self.KP_extent), self.KP_influence, self.aggregation_mode,
                                            deformed, deformed, bias, slope, t_off, t_src,
                                            (ops.grad_dst(self.weights), ops.grad_dst(bias)))
        if deformed:
            self.min_d2 = min_d2
        return out

    def __repr__(self):
        return 'KPConv(radius: {:.2f}, extent: {:.2f}, in_feat: {:d}, out_feat: {:d})'.format(
            self.radius, self.KP_extent, self.in_channels, self.out_channels)


# ----------------------------------------------------------------------------- plain blocks (stock PyTorch)
class BatchNormBlock(nn.Module):
    """BatchNorm1d over stacked points, or a learned bias when batch norm is off (blocks.py:441-478)."""

    def __init__(self, in_dim, use_bn, bn_momentum):
        super().__init__()
        self.bn_momentum = bn_momentum
        self.use_bn = use_bn
        self.in_dim = in_dim
        if use_bn:
            self.batch_norm = nn.BatchNorm1d(in_dim, momentum=bn_momentum)
        else:
            self.bias = Parameter(torch.zeros(in_dim, dtype=torch.float32))

    def reset_parameters(self):
        nn.init.zeros_(self.bias)

    def forward(self, x):
        if not self.use_bn:
            return x + self.bias
        y = self.batch_norm(x.t().unsqueeze(0))  # [1, C, N]
        return y.squeeze(0).t().squeeze()

    def __repr__(self):
        return 'BatchNormBlock(in_feat: {:d}, momentum: {:.3f}, only_bias: {:s})'.format(
            self.in_dim, self.bn_momentum, str(not self.use_bn))


class UnaryBlock(nn.Module):
    """Linear -> BatchNormBlock -> optional LeakyReLU(0.1) (blocks.py:481-515)."""

    def __init__(self, in_dim, out_dim, use_bn, bn_momentum, no_relu=False):
        super().__init__()
        self.bn_momentum = bn_momentum
        self.use_bn = use_bn
        self.no_relu = no_relu
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.mlp = nn.Linear(in_dim, out_dim, bias=True)
        self.batch_norm = BatchNormBlock(out_dim, use_bn, bn_momentum)
        if not no_relu:
            self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, x, batch=None):
        if not self.use_bn and x.is_cuda:
            # Linear + both biases (+ LeakyReLU 0.1) as one tensor-core GEMM with a fused epilogue
            return ops.fused_linear(x, self.mlp.weight, self.mlp.bias, None if self.no_relu else 0.1,
                                    bias2=self.batch_norm.bias)
        x = self.batch_norm(self.mlp(x))
        return x if self.no_relu else self.leaky_relu(x)

    def forward_residual(self, x, residual, slope):
        """leaky_relu(self(x) + residual, slope) for a no_relu block without batch norm, as one GEMM."""
        return ops.fused_linear(x, self.mlp.weight, self.mlp.bias, slope, bias2=self.batch_norm.bias, residual=residual)

    def __repr__(self):
        return 'UnaryBlock(in_feat: {:d}, out_feat: {:d}, BN: {:s}, ReLU: {:s})'.format(
            self.in_dim, self.out_dim, str(self.use_bn), str(not self.no_relu))


class LastUnaryBlock(nn.Module):
    """Bare Linear head (blocks.py:518-541)."""

    def __init__(self, in_dim, out_dim, use_bn, bn_momentum, no_relu=False):
        super().__init__()
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.mlp = nn.Linear(in_dim, out_dim, bias=True)

    def forward(self, x, batch=None):
        if x.is_cuda:
            return ops.fused_linear(x, self.mlp.weight, self.mlp.bias, None)
        return self.mlp(x)

    def __repr__(self):
        return 'LastUnaryBlock(in_feat: {:d}, out_feat: {:d})'.format(self.in_dim, self.out_dim)


def _conv_geometry(block_name, layer_ind, batch):
    """(queries, supports, neighbour matrix) a conv block reads from the collate dict
    (blocks.py:588-595, :660-667): strided blocks go from layer l to l+1 through `pools`."""
    if 'strided' in block_name:
        _ready(batch, ('points', layer_ind), ('points', layer_ind + 1), ('pools', layer_ind))
        return batch['points'][layer_ind + 1], batch['points'][layer_ind], batch['pools'][layer_ind]
    _ready(batch, ('points', layer_ind), ('neighbors', layer_ind))
    return batch['points'][layer_ind], batch['points'][layer_ind], batch['neighbors'][layer_ind]


def _ready(batch, *keys):
    """engine.collate_static builds the pyramid on side streams: wait (on the current stream) for the named levels."""
    pyr = batch.get('_pyramid') if isinstance(batch, dict) else None
    if pyr is not None:
        pyr.wait(*keys)


def _make_kpconv(block_name, in_dim, out_dim, radius, config):
    extent = radius * config.KP_extent / config.conv_radius  # blocks.py:557, :614
    return KPConv(config.num_kernel_points, config.in_points_dim, in_dim, out_dim, extent, radius,
                  fixed_kernel_points=config.fixed_kernel_points, KP_influence=config.KP_influence,
                  aggregation_mode=config.aggregation_mode, deformable='deform' in block_name,
                  modulated=config.modulated)


class SimpleBlock(nn.Module):
    """KPConv -> bias/BN -> LeakyReLU (blocks.py:544-598)."""

    def __init__(self, block_name, in_dim, out_dim, radius, layer_ind, config):
        super().__init__()
        self.bn_momentum = config.batch_norm_momentum
        self.use_bn = config.use_batch_norm
        self.layer_ind = layer_ind
        self.block_name = block_name
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.KPConv = _make_kpconv(block_name, in_dim, out_dim // 2, radius, config)
        self.batch_norm = BatchNormBlock(out_dim // 2, self.use_bn, self.bn_momentum)
        self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, x, batch):
        q_pts, s_pts, inds = _conv_geometry(self.block_name, self.layer_ind, batch)
        if not self.use_bn and x.is_cuda:   # bias + LeakyReLU in the contraction's epilogue
            return self.KPConv(q_pts, s_pts, inds, x, bias=self.batch_norm.bias, slope=0.1)
        return self.leaky_relu(self.batch_norm(self.KPConv(q_pts, s_pts, inds, x)))


class ResnetBottleneckBlock(nn.Module):
    """unary1 -> KPConv -> unary2, plus (max-pooled) shortcut (blocks.py:601-686)."""

    def __init__(self, block_name, in_dim, out_dim, radius, layer_ind, config):
        super().__init__()
        self.bn_momentum = config.batch_norm_momentum
        self.use_bn = config.use_batch_norm
        self.block_name = block_name
        self.layer_ind = layer_ind
        self.in_dim = in_dim
        self.out_dim = out_dim
        mid = out_dim // 4
        self.unary1 = UnaryBlock(in_dim, mid, self.use_bn, self.bn_momentum) if in_dim != mid else nn.Identity()
        self.KPConv = _make_kpconv(block_name, mid, mid, radius, config)
        self.batch_norm_conv = BatchNormBlock(mid, self.use_bn, self.bn_momentum)
        self.unary2 = UnaryBlock(mid, out_dim, self.use_bn, self.bn_momentum, no_relu=True)
        self.unary_shortcut = (UnaryBlock(in_dim, out_dim, self.use_bn, self.bn_momentum, no_relu=True)
                               if in_dim != out_dim else nn.Identity())
        self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, features, batch):
        q_pts, s_pts, inds = _conv_geometry(self.block_name, self.layer_ind, batch)
        strided = 'strided' in self.block_name
        if not self.use_bn and features.is_cuda:
            # conv + bias + LeakyReLU in one op; unary2 + shortcut add + LeakyReLU in one GEMM epilogue.
            # (Running the shortcut as a concurrent forward branch was tried in round 1: autograd then replays its
            # backward on the auxiliary stream, a parameter gradient was lost inside the CUDA-graph capture, and the
            # step gained under 1 % -- the fork/join stays inside the backward nodes, see ops.run_branches.)
            x = self.KPConv(q_pts, s_pts, inds, self.unary1(features), bias=self.batch_norm_conv.bias, slope=0.1)
            sc = self.unary_shortcut(max_pool(features, inds) if strided else features)
            return self.unary2.forward_residual(x, sc, 0.1)
        shortcut = max_pool(features, inds) if strided else features
        x = self.KPConv(q_pts, s_pts, inds, self.unary1(features))
        x = self.unary2(self.leaky_relu(self.batch_norm_conv(x)))
        return self.leaky_relu(x + self.unary_shortcut(shortcut))


class GlobalAverageBlock(nn.Module):
    def forward(self, x, batch):
        return global_average(x, batch['stack_lengths'][-1])


class NearestUpsampleBlock(nn.Module):
    def __init__(self, layer_ind):
        super().__init__()
        self.layer_ind = layer_ind

    def forward(self, x, batch):
        _ready(batch, ('upsamples', self.layer_ind - 1))
        return closest_pool(x, batch['upsamples'][self.layer_ind - 1])

    def __repr__(self):
        return 'NearestUpsampleBlock(layer: {:d} -> {:d})'.format(self.layer_ind, self.layer_ind - 1)


class MaxPoolBlock(nn.Module):
    def __init__(self, layer_ind):
        super().__init__()
        self.layer_ind = layer_ind

    def forward(self, x, batch):
        _ready(batch, ('pools', self.layer_ind + 1))
        return max_pool(x, batch['pools'][self.layer_ind + 1])


_SIMPLE = {'simple' + a + b for a in ('', '_deformable', '_invariant', '_equivariant') for b in ('', '_strided')}
_RESNETB = {'resnetb' + a + b for a in ('', '_deformable', '_invariant', '_equivariant') for b in ('', '_strided')}


def block_decider(block_name, radius, in_dim, out_dim, layer_ind, config):
    """Block factory keyed by the strings of config.architecture (blocks.py:395-438)."""
    if block_name == 'unary':
        return UnaryBlock(in_dim, out_dim, config.use_batch_norm, config.batch_norm_momentum)
    if block_name == 'last_unary':
        return LastUnaryBlock(in_dim, 32, config.use_batch_norm, config.batch_norm_momentum)
    if block_name in _SIMPLE:
        return SimpleBlock(block_name, in_dim, out_dim, radius, layer_ind, config)
    if block_name in _RESNETB:
        return ResnetBottleneckBlock(block_name, in_dim, out_dim, radius, layer_ind, config)
    if block_name in ('max_pool', 'max_pool_wide'):
        return MaxPoolBlock(layer_ind)
    if block_name == 'global_average':
        return GlobalAverageBlock()
    if block_name == 'nearest_upsample':
        return NearestUpsampleBlock(layer_ind)
    raise ValueError('Unknown block name in the architecture definition : ' + block_name)
