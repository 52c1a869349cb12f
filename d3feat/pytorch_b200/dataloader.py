"""Device-side collate -- drop-in for datasets/dataloader.py of the reference.

``collate_fn_descriptor(list_data, config, neighborhood_limits)`` keeps the reference's
signature and returns the same dict (dataloader.py:178-187) with CUDA tensors: the 5-level
point pyramid is built on the GPU by the hash-grid radius-search and grid-subsampling kernels
behind the C ABI, so the CPU cpp_wrappers path (radius_neighbors.batch_query,
grid_subsampling.subsample_batch) disappears.  Inputs may be NumPy arrays (as the reference
dataset yields them), CPU tensors (pinned or not) or CUDA tensors.

CUDA cannot be used in forked DataLoader workers: use ``num_workers=0`` (get_dataloader's
default here) or call the collate in the main process.
"""
from functools import partial

import numpy as np
import torch

from . import ops


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("d3feat.pytorch_b200.dataloader needs a CUDA device (there is no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(a, dtype=None):
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    elif not isinstance(a, torch.Tensor):
        a = torch.as_tensor(a)
    a = a.to(_device(), non_blocking=True)
    return a if dtype is None else a.to(dtype)


def batch_grid_subsampling_kpconv(points, batches_len, features=None, labels=None, sampleDl=0.1, max_p=0, verbose=0,
                                  random_grid_orient=True):
    """Barycentre grid subsampling of stacked clouds (dataloader.py:12-50) -> (s_points, s_len)."""
    if features is not None or labels is not None:
        raise NotImplementedError("feature / label subsampling is not on the D3Feat path (dataloader.py:138 passes "
                                  "points only) and is not implemented by the B200 kernels")
    s_points, s_len = ops.grid_subsample(_to_dev(points, torch.float32), _to_dev(batches_len, torch.int32), sampleDl)
    if max_p > 0:  # per-cloud truncation (grid_subsampling.cpp:174-199)
        lens = s_len.tolist()
        if any(l > max_p for l in lens):
            chunks = [c[:max_p] for c in torch.split(s_points, lens)]
            s_points = torch.cat(chunks, 0)
            s_len = torch.tensor([c.shape[0] for c in chunks], dtype=torch.int32, device=s_points.device)
    return s_points, s_len


def batch_neighbors_kpconv(queries, supports, q_batches, s_batches, radius, max_neighbors, index_dtype=torch.int64):
    """Radius neighbours of stacked clouds (dataloader.py:52-67) -> [Nq, min(max_count, max_neighbors)]."""
    return ops.radius_neighbors(_to_dev(queries, torch.float32), _to_dev(supports, torch.float32),
                                _to_dev(q_batches, torch.int32), _to_dev(s_batches, torch.int32), radius,
                                int(max_neighbors), index_dtype)


def collate_fn_descriptor(list_data, config, neighborhood_limits, index_dtype=torch.int64):
    """Builds the KPFCNN input dict for ONE fragment pair (dataloader.py:69-189)."""
    assert len(list_data) == 1
    pts0, pts1, feat0, feat1, sel_corr, dist_keypts = list_data[0]
    dev = _device()
    pts0, pts1 = _to_dev(pts0, torch.float32), _to_dev(pts1, torch.float32)
    batched_points = torch.cat([pts0, pts1], dim=0)
    batched_features = torch.cat([_to_dev(feat0, torch.float32), _to_dev(feat1, torch.float32)], dim=0)
    batched_lengths = torch.tensor([pts0.shape[0], pts1.shape[0]], dtype=torch.int32, device=dev)

    r_normal = config.first_subsampling_dl * config.conv_radius
    layer_blocks = []
    layer = 0
    input_points, input_neighbors, input_pools, input_upsamples, input_batches_len = [], [], [], [], []
    arch = config.architecture
    empty_idx = torch.zeros((0, 1), dtype=torch.int64, device=dev)
    nb = partial(batch_neighbors_kpconv, index_dtype=index_dtype)

    for block_i, block in enumerate(arch):
        if 'global' in block or 'upsample' in block:
            break
        # accumulate the blocks of the current layer until a pooling/strided block closes it
        if not ('pool' in block or 'strided' in block):
            layer_blocks.append(block)
            if block_i < len(arch) - 1 and 'upsample' not in arch[block_i + 1]:
                continue

        if layer_blocks:
            deform = any('deformable' in b for b in layer_blocks[:-1])
            r = r_normal * config.deform_radius / config.conv_radius if deform else r_normal
            conv_i = nb(batched_points, batched_points, batched_lengths, batched_lengths, r, neighborhood_limits[layer])
        else:
            conv_i = empty_idx

        if 'pool' in block or 'strided' in block:
            dl = 2 * r_normal / config.conv_radius
            pool_p, pool_b = batch_grid_subsampling_kpconv(batched_points, batched_lengths, sampleDl=dl)
            r = r_normal * config.deform_radius / config.conv_radius if 'deformable' in block else r_normal
            pool_i = nb(pool_p, batched_points, pool_b, batched_lengths, r, neighborhood_limits[layer])
            up_i = nb(batched_points, pool_p, batched_lengths, pool_b, 2 * r, neighborhood_limits[layer])
        else:
            pool_i, up_i = empty_idx, empty_idx
            pool_p = torch.zeros((0, 3), dtype=torch.float32, device=dev)
            pool_b = torch.zeros((0,), dtype=torch.int64, device=dev)

        input_points.append(batched_points.float())
        input_neighbors.append(conv_i)
        input_pools.append(pool_i)
        input_upsamples.append(up_i)
        input_batches_len.append(batched_lengths)

        batched_points, batched_lengths = pool_p, pool_b
        r_normal *= 2
        layer += 1
        layer_blocks = []

    return {
        'points': input_points,
        'neighbors': input_neighbors,
        'pools': input_pools,
        'upsamples': input_upsamples,
        'features': batched_features,
        'stack_lengths': input_batches_len,
        'corr': _to_dev(sel_corr),
        'dist_keypts': _to_dev(dist_keypts),
    }


def calibrate_neighbors(dataset, config, collate_fn, keep_ratio=0.8, samples_threshold=2000):
    """Per-layer neighbour limit = `keep_ratio` percentile of the neighbourhood sizes
    (dataloader.py:191-223); the histogram is accumulated on the device."""
    hist_n = int(np.ceil(4 / 3 * np.pi * (config.deform_radius + 1) ** 3))
    hists = None
    for i in range(len(dataset)):
        batch = collate_fn([dataset[i]], config, neighborhood_limits=[hist_n] * config.num_layers)
        rows = []
        for mat in batch['neighbors']:
            counts = (mat < mat.shape[0]).sum(dim=1)
            rows.append(torch.bincount(counts, minlength=hist_n)[:hist_n])
        h = torch.stack(rows)
        hists = h if hists is None else hists + h
        if int(hists.sum(dim=1).min()) > samples_threshold:
            break
    cumsum = torch.cumsum(hists.t(), dim=0)
    percentiles = (cumsum < (keep_ratio * cumsum[hist_n - 1, :])).sum(dim=0)
    return percentiles.cpu().numpy()


def get_dataloader(dataset, batch_size=1, num_workers=0, shuffle=True, neighborhood_limits=None):
    """dataloader.py:225-238.  num_workers must stay 0: the collate launches CUDA kernels."""
    if num_workers != 0:
        raise ValueError("the device-side collate must run in the main process (num_workers=0)")
    if neighborhood_limits is None:
        neighborhood_limits = calibrate_neighbors(dataset, dataset.config, collate_fn=collate_fn_descriptor)
    loader = torch.utils.data.DataLoader(
        dataset, batch_size=batch_size, shuffle=shuffle, num_workers=0,
        collate_fn=partial(collate_fn_descriptor, config=dataset.config, neighborhood_limits=neighborhood_limits),
        drop_last=False)
    return loader, neighborhood_limits
