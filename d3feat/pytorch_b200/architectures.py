"""KPFCNN: the encoder/decoder harness around the KPConv blocks (reference:
models/architectures.py:190-368).  Same constructor, forward contract
``model(batch) -> (features [N,32] L2-normalised, scores [N,1])`` and state_dict keys;
the KPConv layers inside run on the sm_100a kernels, everything else is stock PyTorch.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

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
        if verbose:
            print(self)

    def forward(self, batch):
        x = batch['features'].clone().detach()
        skips = []
        for i, block in enumerate(self.encoder_blocks):
            if i in self.encoder_skips:
                skips.append(x)
            x = block(x, batch)
        for j, block in enumerate(self.decoder_blocks):
            if j in self.decoder_concats:
                x = torch.cat([x, skips.pop()], dim=1)
            x = block(x, batch)
        scores = self.detection_scores(batch, x)
        return F.normalize(x, p=2, dim=-1), scores

    def detection_scores(self, inputs, features):
        """Saliency x channel-max keypoint score (architectures.py:322-368); stock PyTorch
        (next in line for a fused kernel)."""
        neighbor = inputs['neighbors'][0].long()
        n = features.shape[0]
        feats = torch.cat([features, torch.zeros_like(features[:1])], dim=0)
        neighbor = torch.cat([neighbor, torch.full_like(neighbor[:1], n)], dim=0)
        feats = feats / (feats.max() + 1e-6)
        nf = feats[neighbor]                                            # [N+1, H, C]
        count = (nf.sum(dim=-1) != 0).sum(dim=-1, keepdim=True).clamp(min=1)
        local_max_score = F.softplus(feats - nf.sum(dim=1) / count)
        depth_wise_max_score = feats / (1e-6 + feats.max(dim=1, keepdim=True)[0])
        scores = (local_max_score * depth_wise_max_score).max(dim=1, keepdim=True)[0]
        if not self.training:  # hard local-max gate at test time
            is_local_max = feats == nf.max(dim=1)[0]
            scores = scores * is_local_max.float().max(dim=1, keepdim=True)[0]
        return scores[:-1]
