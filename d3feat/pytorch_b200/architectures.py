"""KPFCNN: the encoder/decoder harness around the KPConv blocks (reference:
models/architectures.py:190-368).  Same constructor, forward contract
``model(batch) -> (features [N,32] L2-normalised, scores [N,1])`` and state_dict keys;
the KPConv layers inside run on the sm_100a kernels, everything else is stock PyTorch.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .blocks import block_decider

_LAYER_CHANGE = ('pool', 'strided', 'upsample', 'global')


class KPFCNN(nn.Module):

    def __init__(self, config, verbose=False):
        super().__init__()
        arch = list(config.architecture)
        self.K = config.num_kernel_points
        layer = 0
        r = config.first_subsampling_dl * config.conv_radius
        in_dim, out_dim = config.in_features_dim, config.first_features_dim

        # ---- encoder: every block up to the first upsampling (architectures.py:213-250)
        self.encoder_blocks = nn.ModuleList()
        self.encoder_skip_dims = []
        self.encoder_skips = []
        for i, name in enumerate(arch):
            if 'equivariant' in name and out_dim % 3 != 0:
                raise ValueError('Equivariant block but features dimension is not a factor of 3')
            if any(tag in name for tag in _LAYER_CHANGE):
                self.encoder_skips.append(i)
                self.encoder_skip_dims.append(in_dim)
            if 'upsample' in name:
                break
            self.encoder_blocks.append(block_decider(name, r, in_dim, out_dim, layer, config))
            in_dim = out_dim // 2 if 'simple' in name else out_dim
            if 'pool' in name or 'strided' in name:
                layer, r, out_dim = layer + 1, r * 2, out_dim * 2

        # ---- decoder: from the first upsampling on (architectures.py:252-294)
        self.decoder_blocks = nn.ModuleList()
        self.decoder_concats = []
        first_up = next((i for i, name in enumerate(arch) if 'upsample' in name), 0)
        for j, name in enumerate(arch[first_up:]):
            if j > 0 and 'upsample' in arch[first_up + j - 1]:
                in_dim += self.encoder_skip_dims[layer]
                self.decoder_concats.append(j)
            self.decoder_blocks.append(block_decider(name, r, in_dim, out_dim, layer, config))
            in_dim = out_dim
            if 'upsample' in name:
                layer, r, out_dim = layer - 1, r * 0.5, out_dim // 2
        # engine.PairStep may install a callback that fires during backward as soon as the gradients of the decoder and
        # of the encoder blocks from `_early_block` on are complete (first bucket of the data-parallel all-reduce)
        self._early_block = None
        self._on_early_grads = None
        if verbose:
            print(self)

    def _early_hook(self, grad):
        if self._on_early_grads is not None:
            self._on_early_grads()
        return None

    def forward(self, batch):
        x = batch['features'].clone().detach()
        skips = []
        for i, block in enumerate(self.encoder_blocks):
            if i in self.encoder_skips:
                skips.append(x)
            if i == self._early_block and self._on_early_grads is not None and x.requires_grad:
                x.register_hook(self._early_hook)
            x = block(x, batch)
        for j, block in enumerate(self.decoder_blocks):
            if j in self.decoder_concats:
                x = torch.cat([x, skips.pop()], dim=1)
            x = block(x, batch)
        scores = self.detection_scores(batch, x)
        return F.normalize(x, p=2, dim=-1), scores

    def detection_scores(self, inputs, features):
        """Saliency x channel-max keypoint score (architectures.py:322-368): neighbour-mean saliency
        softplus(f - mean_nb f), depth-wise ratio f / max_c f, max over channels, and at test time the
        exact-equality local-max gate.  One fused warp-per-point kernel (d3f_detection_scores_*)."""
        pyr = inputs.get('_pyramid') if isinstance(inputs, dict) else None
        if pyr is not None:   # engine.collate_static builds the pyramid on side streams
            pyr.wait(('neighbors', 0))
        return ops.detection_scores(features, inputs['neighbors'][0], not self.training)
