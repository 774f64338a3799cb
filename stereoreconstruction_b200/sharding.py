"""Multi-GPU partitioning of the dense-matching path.

Every reference pixel is independent (stereo/multiviewstereo.cpp:543-604), inputs are replicated
on every rank, so the data path needs NO collective: reference views (multi-view runs) or row
bands (two-view / large images) are dealt to ranks, each rank runs sr_run_view on its share, and
the depth/index maps are gathered once at the end (cross-check reads every view's depths,
multiviewstereo.cpp:694-719).  This module is the host-side bookkeeping shared by bench.py and
the tests; the gather uses torch.distributed (NCCL on GPUs, gloo in the CPU tests) or, inside the
library, sr_comm_allgather_views / sr_comm_allgather_rows (NCCL broadcasts from the owners).
"""
import numpy as np


def partition_views(num_views, world_size):
    """views[r] = reference views owned by rank r: round-robin, so that neighbouring (similar-cost)
    views land on different ranks."""
    return [[v for v in range(num_views) if v % world_size == r] for r in range(world_size)]


def view_owner(num_views, world_size):
    """owner[v] = rank that computes reference view v (the argument of sr_comm_allgather_views)."""
    return np.array([v % world_size for v in range(num_views)], dtype=np.int32)


def row_bands(height, world_size, align=1):
    """[(row_begin, row_end)] per rank: contiguous, covering [0, height), sizes differing by at most
    `align` rows; a rank may get an empty band when height < world_size."""
    units = (height + align - 1) // align
    bands, start = [], 0
    for r in range(world_size):
        n = units // world_size + (1 if r < units % world_size else 0)
        b0, b1 = min(height, start * align), min(height, (start + n) * align)
        bands.append((b0, b1))
        start += n
    return bands


def plan(num_views, height, world_size):
    """Work list per rank as (view, row_begin, row_end).  With at least as many views as ranks the
    unit is a whole view; with fewer (two-view runs, or more GPUs than views) every view is split
    into row bands over all ranks ('for G > V combine with row sharding')."""
    if num_views >= world_size:
        return [[(v, 0, height) for v in vs] for vs in partition_views(num_views, world_size)]
    bands = row_bands(height, world_size)
    return [[(v, b0, b1) for v in range(num_views) if b1 > b0] for (b0, b1) in bands]


def gather_maps(local, work, num_views, shape, dist=None, fill=None):
    """All-gathers per-view maps.  local[(v, b0, b1)] = array of rows [b0, b1) of view v computed by
    this rank; work = plan(...) for all ranks; returns the list of full (h, w) maps on every rank.
    `dist` is torch.distributed (initialised) or None for a single process."""
    h, w = shape
    sample = next(iter(local.values())) if local else None
    dtype = sample.dtype if sample is not None else np.float64
    full = [np.full((h, w), fill if fill is not None else 0, dtype=dtype) for _ in range(num_views)]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        for (v, b0, b1), a in local.items():
            full[v][b0:b1] = a
        return full
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    for r in range(world):  # one broadcast per work item from its owner: each byte moves once
        for (v, b0, b1) in work[r]:
            if r == rank:
                t = torch.from_numpy(np.ascontiguousarray(local[(v, b0, b1)])).to(dev)
            else:
                t = torch.empty((b1 - b0, w), dtype=torch.from_numpy(np.empty(0, dtype=dtype)).dtype, device=dev)
            dist.broadcast(t, src=r)
            full[v][b0:b1] = t.cpu().numpy()
    return full
