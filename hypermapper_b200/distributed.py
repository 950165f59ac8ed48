"""Multi-GPU sharding of the candidate axis (one process per GPU, torch.distributed).

The candidate pool shards naturally: rank r scores global indices [r*N/G, (r+1)*N/G) with a replicated model
and no data-path collective.  The only exchange is the per-rank stable top-k: k (value, global index) pairs per
rank are all-gathered (k*16 bytes per rank -- latency, not bandwidth) and every rank merges the G*k pairs with the
same (value, index) ordering, so all ranks hold the identical global top-k, equal to the single-GPU result and to
the reference's first-index tie-break (np.argmin).  Backend: nccl on GPUs, gloo in the CPU tests.
"""
import numpy as np


def shard_bounds(n_total, rank, world):
    """contiguous slice of the candidate axis owned by `rank`"""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi


def merge_topk_host(values, indices, k):
    """stable merge of gathered (value, global index) pairs on the host: NaN first, -0.0 == +0.0, ties -> lowest
    global index; pads (+inf, -1) are dropped."""
    v = np.asarray(values, dtype=np.float64).reshape(-1)
    i = np.asarray(indices, dtype=np.int64).reshape(-1)
    keep = i >= 0
    v, i = v[keep], i[keep]
    nan = np.isnan(v)
    key = np.where(nan, -np.inf, v + 0.0)
    order = np.lexsort((i, key, ~nan))[:k]
    return v[order], i[order]


def allgather_topk(local_vals, local_idx, k, group=None):
    """local_vals/local_idx: tensors [k] (CUDA with nccl, CPU with gloo) holding this rank's top-k with GLOBAL indices.
    Returns the global top-k (values[k], indices[k]) as tensors on the same device, identical on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local_vals, local_idx
    gv = torch.empty((world * local_vals.numel(),), dtype=local_vals.dtype, device=local_vals.device)
    gi = torch.empty((world * local_idx.numel(),), dtype=local_idx.dtype, device=local_idx.device)
    dist.all_gather_into_tensor(gv, local_vals.contiguous().view(-1), group=group)
    dist.all_gather_into_tensor(gi, local_idx.contiguous().view(-1), group=group)
    if gv.is_cuda:
        from .forest import topk_merge
        return topk_merge(gv, gi, k)
    mv, mi = merge_topk_host(gv.numpy(), gi.numpy(), k)
    out_v = torch.full((k,), float("inf"), dtype=torch.float64)
    out_i = torch.full((k,), -1, dtype=torch.int64)
    out_v[: len(mv)] = torch.from_numpy(mv)
    out_i[: len(mi)] = torch.from_numpy(mi)
    return out_v, out_i
