"""Multi-GPU plumbing: the path shards by sample — one process per GPU, weights replicated,
no collective on the data path.  The only exchange is the end-of-batch gather of (epe_sum: f64, n: i64) per
rank, mirroring the EPE metric state of utils/metrics.py:30-49 and ``epe_masked`` (utils/metrics.py:196-213).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of the global batch owned by ``rank`` (contiguous, remainder to the low ranks)."""
    assert 0 <= rank < world and global_batch >= 0
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t: Optional[torch.Tensor], rank: int, world: int) -> Optional[torch.Tensor]:
    if t is None:
        return None
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def epe_sum_count(flow: torch.Tensor, target: torch.Tensor, valid: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-pixel end-point error sqrt(sum_c (flow-target)^2), masked (utils/metrics.py:196-213) → (sum f64, count i64)."""
    if flow.is_cuda:
        from .events import epe_sum_count as _device_epe          # bflow_epe_masked kernel
        return _device_epe(flow, target, valid)
    e = torch.sqrt(((flow - target) ** 2).sum(dim=1))
    if valid is not None:
        v = valid.reshape(e.shape).bool()
        return e[v].double().sum(), v.sum().to(torch.int64)
    return e.double().sum(), torch.tensor(e.numel(), dtype=torch.int64, device=e.device)


def gather_epe(epe_sum: torch.Tensor, count: torch.Tensor) -> Tuple[float, int, torch.Tensor]:
    """all_gather of the 16-byte (sum, count) state of every rank; returns (global mean EPE, global count,
    per-rank table (world, 2) as float64).  Works on NCCL (CUDA tensors) and gloo (CPU tensors)."""
    state = torch.stack([epe_sum.double().reshape(()), count.double().reshape(())])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        world = dist.get_world_size()
        table = torch.empty(world, 2, dtype=torch.float64, device=state.device)
        dist.all_gather_into_tensor(table, state) if state.is_cuda else dist.all_gather(list(table.unbind(0)), state)
    else:
        table = state[None]
    tot, n = table[:, 0].sum().item(), int(table[:, 1].sum().item())
    return (tot / n if n else float('nan')), n, table
