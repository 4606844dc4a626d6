"""Multi-GPU plumbing for the scoring path: one process per GPU, videos sharded by rank, no data-path
collective; the only exchange is the final gather of per-video logits (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  Replaces the reference's single-process ``nn.DataParallel`` scatter/gather over dim 0
(train.py:249-251) for inference; the per-video mean stays rank-local because a video's frames never leave its GPU.
"""
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_range(num_items: int, rank: int, world: int):
    """Contiguous, balanced block of ``[0, num_items)`` owned by ``rank`` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(num_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_counts(num_items: int, world: int) -> List[int]:
    return [shard_range(num_items, r, world)[1] - shard_range(num_items, r, world)[0] for r in range(world)]


def gather_logits(local: torch.Tensor, num_items: int, group=None) -> torch.Tensor:
    """All ranks contribute their block of per-video logits ([n_local] or [n_local, 1]); every rank gets the
    full ``[num_items]`` vector in the original video order.  Ragged blocks are padded to the largest one so a
    single fixed-size all_gather suffices (logits are a few hundred bytes: latency-bound, not bandwidth-bound)."""
    if not (dist.is_available() and dist.is_initialized()):
        out = local.reshape(-1)
        if out.numel() != num_items:
            raise ValueError("single-process gather: local block must hold every item")
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = shard_counts(num_items, world)
    flat = local.reshape(-1).float()
    if flat.numel() != counts[rank]:
        raise ValueError(f"rank {rank} holds {flat.numel()} logits, its shard has {counts[rank]}")
    width = max(max(counts), 1)
    send = torch.zeros(width, dtype=torch.float32, device=flat.device)
    send[: flat.numel()] = flat
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return torch.cat([recv[r][: counts[r]] for r in range(world)])


def score_videos_sharded(score_fn: Callable[[torch.Tensor], torch.Tensor], make_videos: Callable[[Sequence[int]], torch.Tensor],
                         num_videos: int, videos_per_call: int, group=None, device: Optional[torch.device] = None):
    """Eval-style scoring (reference eval.py:135-194 with the DataLoader replaced by ``make_videos``):
    rank r scores videos ``shard_range(num_videos, r, world)`` in calls of ``videos_per_call`` videos through
    ``score_fn(x[B,K,C,H,W]) -> logits[B] or [B,1]`` and the logits are gathered on every rank."""
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_range(num_videos, rank, world)
    parts = []
    for s in range(lo, hi, videos_per_call):
        ids = list(range(s, min(s + videos_per_call, hi)))
        parts.append(score_fn(make_videos(ids)).reshape(-1).float())
    if parts:
        local = torch.cat(parts)
    else:
        local = torch.zeros(0, dtype=torch.float32, device=device or "cpu")
    return gather_logits(local, num_videos, group)
