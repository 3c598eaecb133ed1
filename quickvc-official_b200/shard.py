"""Utterance-level sharding across the GPUs of one box.

`SynthesizerTrn.infer` has no cross-utterance operation (models.py:625-642), so a batch of utterances is
split contiguously over the ranks, every rank converts its slice with its own replica of the folded
weights, and the only exchange is the final gather of the waveforms to rank 0 -- no collective on the
hot path.  One process per GPU, `torch.distributed` for the plumbing (NCCL on the box, gloo in the CPU
tests); the per-rank conversion is whatever callable the caller passes (the B200 module's `infer`).
"""
from __future__ import annotations

import os
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


class HostGather:
    """The final gather done where the result has to end up anyway: ONE host buffer, shared by the ranks of the box and
    page-locked in each of them, into which every rank copies its own waveforms over its own PCIe link while it
    converts the next utterances.  No device-to-device traffic, no rank-0 bottleneck (a `dist.gather` to rank 0 followed by
    one 2.6 GB device-to-host copy there costs as much as converting 4096 utterances on 8 GPUs).

    The buffer is a file in /dev/shm mapped by all ranks; rank 0 creates and removes it."""

    def __init__(self, n: int, samples: int, *, name: Optional[str] = None, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n, self.samples = n, samples
        tag = name or f"qvc_gather_{os.environ.get('MASTER_PORT', str(os.getpid()))}_{n}x{samples}"
        self.path = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", tag)
        numel = max(1, n * samples)
        if self.rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(numel * 4)
        if self.world > 1:
            dist.barrier(group=group)
        self.flat = torch.from_file(self.path, shared=True, size=numel, dtype=torch.float32)
        self.host = self.flat[: n * samples].view(n, 1, samples)
        # page-locked in every rank through the library (a failure there -- e.g. a locked-memory limit -- leaves the copies
        # correct, only synchronous, and no error pending in torch's CUDA runtime)
        self._registered = False
        if torch.cuda.is_available():
            from . import capi
            torch.cuda.current_device()                # make sure this process has a CUDA context
            self._registered = capi.load().qvc_host_register(self.flat.data_ptr(), numel * 4) == 0
        self._copy_stream = None

    def close(self) -> None:
        if self._registered:
            from . import capi
            torch.cuda.synchronize()
            capi.load().qvc_host_unregister(self.flat.data_ptr())
            self._registered = False
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.rank == 0 and os.path.exists(self.path):
            os.unlink(self.path)

    def convert(self, infer: Callable[[Tensor, Tensor], Tensor], unit: Tensor, mel: Tensor, *, chunk: int = 64) -> Optional[Tensor]:
        """Every rank converts its contiguous slice of `unit` (N, 256, T), `chunk` utterances per `infer` call, and copies each
        chunk's waveforms into its rows of the shared host buffer on a side stream while the next chunk runs.  On return the
        calling stream has waited for this rank's copies; after the closing barrier rank 0 returns the (N, 1, 320 T) host
        tensor holding every rank's rows, the other ranks None."""
        n, _, t = unit.shape
        if n != self.n or 320 * t != self.samples:
            raise ValueError(f"HostGather built for ({self.n}, {self.samples}), got ({n}, {320 * t})")
        lo, hi = shard_range(n, self.world, self.rank)
        on_gpu = unit.is_cuda
        if on_gpu and self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=unit.device)
        for i in range(lo, hi, chunk):
            j = min(hi, i + chunk)
            out = infer(unit[i:j], mel)
            if out.shape != (j - i, 1, 320 * t):
                raise RuntimeError(f"infer returned {tuple(out.shape)} for a slice of {j - i} utterances of {t} frames")
            if on_gpu:
                self._copy_stream.wait_stream(torch.cuda.current_stream(unit.device))
                with torch.cuda.stream(self._copy_stream):
                    self.host[i:j].copy_(out, non_blocking=True)
                out.record_stream(self._copy_stream)
            else:
                self.host[i:j].copy_(out)
        if on_gpu:
            torch.cuda.current_stream(unit.device).wait_stream(self._copy_stream)
        return self.host if self.rank == 0 else None

    def finish(self, device=None) -> None:
        """Blocks until this rank's copies have landed and every rank has reached this point: rank 0 may read the buffer."""
        if device is not None and torch.cuda.is_available():
            torch.cuda.synchronize(device)
        if self.world > 1:
            dist.barrier(group=self.group)
