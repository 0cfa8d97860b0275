"""Multi-GPU sharding of the hot path: one process per GPU, contiguous shards, NO collective on the
compute path (contracts are independent — _model_calls.py:84, _iv_models.py:29 loop bodies).

`torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests) is used only for the optional
steps around it:
  * gather of the output columns to rank 0 when the caller wants one result,
  * the global view of the per-contract status bytes (ascending index lists for the reference's
    "Found Below Intrinsic contracts at index [...]" warnings, util/data_format.py:56-68),
  * the min-reduction of the first ZeroDivisionError index.
"""
import numpy as np


def shard_bounds(n, world, rank):
    """Contiguous range [lo, hi) of rank `rank`: keeps strike-sorted chains warp-coherent."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def shard_columns(cols, n, world, rank):
    """Slice dense columns to this rank's shard; broadcast scalars (size 1) pass through."""
    lo, hi = shard_bounds(n, world, rank)
    out = {}
    for k, v in cols.items():
        if v is None:
            out[k] = None
        elif np.size(v) == 1 and n != 1:
            out[k] = v
        else:
            out[k] = v[lo:hi]
    return out, lo, hi


def _dist():
    import torch.distributed as dist
    return dist


def gather_column(local, n, dst=0, group=None):
    """Gather a 1-D torch tensor shard (sizes may differ by one) to rank `dst`; returns the full
    column on dst, None elsewhere."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
    m = max(sizes)
    pad = local
    if local.numel() < m:
        pad = torch.empty(m, dtype=local.dtype, device=local.device)
        pad[:local.numel()] = local
    bufs = [torch.empty(m, dtype=local.dtype, device=local.device) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def global_status_indices(status_local, lo, bit, group=None):
    """Ascending global indices (as np.argwhere gives them in the reference) of the contracts whose
    status byte has `bit` set; every rank receives the full list."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    idx = torch.nonzero(status_local & bit, as_tuple=False).ravel().to(torch.int64) + lo
    cnt = torch.tensor([idx.numel()], dtype=torch.int64, device=status_local.device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    m = int(max(c.item() for c in cnts))
    if m == 0:
        return []
    pad = torch.full((m,), -1, dtype=torch.int64, device=status_local.device)
    pad[:idx.numel()] = idx
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = []
    for b, c in zip(bufs, cnts):
        out.extend(b[:int(c.item())].tolist())
    return out


def reduce_first_zde(first_local, group=None, device=None):
    """first_local: global index of this shard's first zero-division (or -1).  Returns the global first."""
    import torch
    dist = _dist()
    big = np.iinfo(np.int64).max
    t = torch.tensor([big if first_local < 0 else first_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    v = int(t.item())
    return -1 if v == big else v
