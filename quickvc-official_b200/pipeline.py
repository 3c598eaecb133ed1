"""Host <-> device pipelining for bulk conversion (the serving shape of `convert.py:59-86`).

`convert.py` moves each utterance to the GPU, calls `infer`, and copies the waveform back, one after the
other.  For batches that keeps the GPU idle during both PCIe transfers (33 MB in, 41 MB out for 64 x 10 s).
`PipelinedConverter` runs the same three steps -- H2D of `unit` / `mel`, `SynthesizerTrn.infer`, D2H of the
waveform -- on three CUDA streams with double-buffered device inputs and pinned host outputs, so the copies of
batch i-1 and i+1 overlap the kernels of batch i.  The arithmetic is exactly `net.infer(unit, mel)`.
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Tuple

import torch
from torch import Tensor


class GraphedInfer:
    """One `infer` call of fixed shape captured into a CUDA graph and replayed.

    The ~110 kernel launches of a call (and the fork / join of the speaker-encoder side stream) become one graph
    launch: the GPU-side time is unchanged -- the chain of dependent kernels is what a single clip costs
    (profiles/r01_final_summary.md) -- but the host thread spends ~20 us per call instead of ~1 ms, which is what a
    server multiplexing many streams needs.  Inputs are copied into static buffers; the returned waveform is the
    graph's static output buffer (valid until the next call; clone it to keep it).
    """

    def __init__(self, net, batch: int, frames: int, *, mel_frames: int = 0, device: Optional[torch.device] = None) -> None:
        """mel_frames > 0: `infer(unit, mel)` with a (1, 80, mel_frames) target mel; 0: `infer_with_embedding(unit, g)`
        with `g` (1 | batch, 256) given at the first call (its batch size is then fixed)."""
        self.net = net
        self.device = device or next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("GraphedInfer needs the module on a CUDA device")
        d = self.device
        self._unit = torch.zeros(batch, 256, frames, device=d)
        self._noise = torch.zeros(batch, 192, frames, device=d)
        self._mel = torch.zeros(1, 80, mel_frames, device=d) if mel_frames > 0 else None
        self._g: Optional[Tensor] = None
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._wave: Optional[Tensor] = None

    def _run(self) -> Tensor:
        if self._mel is not None:
            return self.net.infer(self._unit, self._mel, noise=self._noise)
        return self.net.infer_with_embedding(self._unit, self._g, noise=self._noise)

    def _capture(self) -> None:
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):           # warm-up outside the capture: weight folding, workspaces, attributes
            self._run()
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._wave = self._run()

    @torch.no_grad()
    def __call__(self, unit: Tensor, cond: Tensor, noise: Optional[Tensor] = None) -> Tensor:
        """`cond` is the target mel (mel_frames > 0) or the speaker embedding(s)."""
        self._unit.copy_(unit, non_blocking=True)
        if noise is None:
            self._noise.normal_()                                   # models.py:94
        else:
            self._noise.copy_(noise, non_blocking=True)
        if self._mel is not None:
            self._mel.copy_(cond, non_blocking=True)
        else:
            cond = cond.reshape(cond.shape[0], -1)
            if self._g is None:
                self._g = torch.zeros_like(cond, device=self.device)
            self._g.copy_(cond, non_blocking=True)
        if self._graph is None:
            self._capture()
        self._graph.replay()
        return self._wave


class PipelinedConverter:
    """Converts a stream of host batches `(unit (B,256,T), mel (1,80,Tm))` to host waveforms `(B,1,320 T)`.

    All batches must share one shape (pad or bucket upstream).  Inputs should live in pinned memory for the
    copies to be asynchronous; outputs are written into pinned buffers owned by the converter (`depth` of
    them: a result stays valid until `depth` further batches have been converted) or into `out` buffers
    supplied per call.
    """

    def __init__(self, net, batch: int, frames: int, mel_frames: int, *, device: Optional[torch.device] = None,
                 depth: int = 2) -> None:
        self.net = net
        self.device = device or next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("PipelinedConverter needs the module on a CUDA device")
        self.depth = depth
        d = self.device
        self._unit = [torch.empty(batch, 256, frames, device=d) for _ in range(depth)]
        self._mel = [torch.empty(1, 80, mel_frames, device=d) for _ in range(depth)]
        self._wave_h = [torch.empty(batch, 1, 320 * frames).pin_memory() for _ in range(depth)]
        self._s_in, self._s_out = torch.cuda.Stream(d), torch.cuda.Stream(d)
        self._ev_in = [torch.cuda.Event() for _ in range(depth)]
        self._ev_done = [torch.cuda.Event() for _ in range(depth)]     # compute of slot finished (inputs reusable)
        self._ev_out = [torch.cuda.Event() for _ in range(depth)]      # D2H of slot finished (host buffer valid)
        self._n = 0

    def submit(self, unit_h: Tensor, mel_h: Tensor, noise: Optional[Tensor] = None) -> Tuple[Tensor, torch.cuda.Event]:
        """Enqueue one batch; returns (pinned host waveform buffer, event that fires when it is filled)."""
        k = self._n % self.depth
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._s_in):
            if self._n >= self.depth:
                self._s_in.wait_event(self._ev_done[k])          # the kernels that read this slot are done
            self._unit[k].copy_(unit_h, non_blocking=True)
            self._mel[k].copy_(mel_h, non_blocking=True)
            self._ev_in[k].record(self._s_in)
        main.wait_event(self._ev_in[k])
        wave = self.net.infer(self._unit[k], self._mel[k], noise=noise)
        self._ev_done[k].record(main)
        with torch.cuda.stream(self._s_out):
            self._s_out.wait_event(self._ev_done[k])
            if self._n >= self.depth:
                pass                                             # host buffer k is overwritten: see class docstring
            self._wave_h[k].copy_(wave, non_blocking=True)
            wave.record_stream(self._s_out)                      # keep the allocation alive until the copy is done
            self._ev_out[k].record(self._s_out)
        self._n += 1
        return self._wave_h[k], self._ev_out[k]

    def drain(self) -> None:
        """Make the calling stream wait for every outstanding copy (call before timing stops / reading results)."""
        main = torch.cuda.current_stream(self.device)
        for e in self._ev_out[: min(self._n, self.depth)]:
            main.wait_event(e)

    def convert_many(self, batches: Iterable[Tuple[Tensor, Tensor]]) -> Iterator[Tensor]:
        """Yields one host waveform tensor (a copy) per input batch, in order."""
        pending: List[Tuple[Tensor, torch.cuda.Event]] = []
        for unit_h, mel_h in batches:
            pending.append(self.submit(unit_h, mel_h))
            if len(pending) == self.depth:
                buf, ev = pending.pop(0)
                ev.synchronize()
                yield buf.clone()
        for buf, ev in pending:
            ev.synchronize()
            yield buf.clone()
