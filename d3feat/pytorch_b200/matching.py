"""Test-time descriptor matching -- drop-in for geometric_registration/common.py::build_correspondence of the
reference.  The reference forms the dense N x M distance matrix in NumPy on the host for every
fragment pair; here one warp per keypoint scans the other set on the GPU (d3f_mutual_nn) and the matrix is never stored.
There is no CPU path: the descriptors are moved to the current CUDA device."""
import numpy as np
import torch

from . import ops


def build_correspondence(source_desc, target_desc):
    """Mutually closest pairs in feature space (reference signature and return value: an int array of
    [source index, target index] rows in ascending source index; an empty result has shape (0,) like np.array([]))."""
    if not torch.cuda.is_available():
        raise RuntimeError("d3feat.pytorch_b200: build_correspondence needs a CUDA device (there is no CPU path)")
    dev = torch.device("cuda", torch.cuda.current_device())
    s = torch.as_tensor(source_desc, dtype=torch.float32).to(dev)
    t = torch.as_tensor(target_desc, dtype=torch.float32).to(dev)
    pairs, _, _ = ops.mutual_nn(s, t)
    out = pairs.cpu().numpy().astype(np.int64)
    return out if out.shape[0] else np.array([])
