"""Multi-GPU plumbing of the rollout: instances are independent, so they are sharded over ranks with no
collective on the decode path; the only exchange is one gather of the [B_local] fp32 tour costs
(SURVEY.md section 8e).  One process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_instances: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) instance range of ``rank``; the first n % world ranks hold one extra instance."""
    base, rem = divmod(n_instances, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_costs(local_cost: torch.Tensor, n_instances: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather the per-rank cost vectors (ragged: shard sizes differ by at most one) to ``dst`` in instance order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_cost
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_instances, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(width, dtype=local_cost.dtype, device=local_cost.device)
    buf[: local_cost.numel()] = local_cost
    out: Optional[List[torch.Tensor]] = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)])


def sharded_rollout(engine, data: torch.Tensor, decode_type: str = "greedy", dst: int = 0, **kw):
    """Every rank holds (or is handed) the full batch; it decodes its shard and rank ``dst`` receives all costs."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(data.shape[0], rank, world)
    out = engine.rollout(data[lo:hi], decode_type, **kw)
    return out, gather_costs(out.R, data.shape[0], dst)
