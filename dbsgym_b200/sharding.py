"""Multi-GPU: environments are independent (no cross-env term in reference env.py:252-256), so
they shard across ranks as contiguous index ranges with NO collective on the step path.  The only
exchange is a gather of per-episode statistics at episode boundaries (torch.distributed: NCCL on
GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_envs: int, rank: int, world_size: int):
    """Contiguous [lo, hi) slice of the global env index range owned by ``rank`` (sizes differ by <= 1)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(n_envs, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_params(params_dicts, rank: int, world_size: int):
    lo, hi = shard_bounds(len(params_dicts), rank, world_size)
    return params_dicts[lo:hi], lo


def gather_episode_stats(local_stats, n_envs_global: int, rank: int, world_size: int, device=None):
    """All-gather a [n_local, k] float64 table of per-env episode statistics into the global
    [n_envs_global, k] table (identical on every rank).  Shards may differ in size by one row, so
    rows are padded to the largest shard before the collective."""
    import torch
    import torch.distributed as dist
    local = torch.as_tensor(np.asarray(local_stats, dtype=np.float64))
    if device is not None:
        local = local.to(device)
    k = local.shape[1]
    sizes = [shard_bounds(n_envs_global, r, world_size) for r in range(world_size)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((pad, k), dtype=torch.float64, device=local.device)
    buf[: local.shape[0]] = local
    if world_size == 1 or not (dist.is_available() and dist.is_initialized()):
        parts = [buf]
    else:
        parts = [torch.empty_like(buf) for _ in range(world_size)]
        dist.all_gather(parts, buf)
    out = torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
    return out.cpu().numpy()
