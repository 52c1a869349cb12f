"""d3feat.pytorch_b200 -- B200 (sm_100a) implementation of D3Feat's data-parallel hot path.

Host code mirrors the reference's Python surface (KPConv / blocks / KPFCNN, the
collate_fn_descriptor dict API, CircleLoss / ContrastiveLoss / DetLoss); the work is done by
hand-written CUDA kernels in libd3feat_b200.so behind the C ABI of include/d3feat_b200.h.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops", "blocks", "architectures", "dataloader", "loss", "config", "synthetic", "kernel_points"]
