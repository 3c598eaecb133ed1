"""Utterance-level sharding across the GPUs of one box.

`SynthesizerTrn.infer` has no cross-utterance operation (models.py:625-642), so a batch of utterances is
split contiguously over the ranks, every rank converts its slice with its own replica of the folded
weights, and the only exchange is the final gather of the waveforms to rank 0 -- no collective on the
hot path.  One process per GPU, `torch.distributed` for the plumbing (NCCL on the box, gloo in the CPU
tests); the per-rank conversion is whatever callable the caller passes (the B200 module's `infer`).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n utterances owned by `rank`: the first n % world ranks get one more."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def convert_sharded(infer: Callable[[Tensor, Tensor], Tensor], unit: Tensor, mel: Tensor, *,
                    group: Optional[dist.ProcessGroup] = None, gather: bool = True) -> Optional[Tensor]:
    """Convert `unit` (N, 256, T) -- the same full batch on every rank -- to the target speaker `mel`.

    Each rank runs `infer(unit[lo:hi], mel)` on its slice.  With `gather`, rank 0 returns the (N, 1, 320 T)
    waveforms in utterance order and the other ranks return None; without it every rank returns its slice.
    Ranks with an empty slice (N < world) contribute nothing.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n, _, t = unit.shape
    lo, hi = shard_range(n, world, rank)
    mine = infer(unit[lo:hi], mel) if hi > lo else unit.new_zeros((0, 1, 320 * t))
    if mine.shape != (hi - lo, 1, 320 * t):
        raise RuntimeError(f"infer returned {tuple(mine.shape)} for a slice of {hi - lo} utterances of {t} frames")
    if not gather:
        return mine
    if world == 1:
        return mine
    # equal-size gather: pad every slice to the largest one (at most one utterance of padding)
    cap = -(-n // world)
    buf = mine.new_zeros((cap, 1, 320 * t))
    buf[: hi - lo] = mine
    out: Optional[List[Tensor]] = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0, group=group)
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        rlo, rhi = shard_range(n, world, r)
        parts.append(out[r][: rhi - rlo])
    return torch.cat(parts, 0)
