"""Host-side plumbing between `SynthesizerTrn` and libqvc_b200.so: folded-weight cache, workspace
cache, argument checking, pointer marshalling.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import capi

_PRECISIONS = {"fp32": capi.OPF_F32, "tf32": capi.OPF_TF32, "bf16": capi.OPF_BF16, "fp16": capi.OPF_F16}
_BACKENDS = {"fma": capi.BACKEND_FMA, "tcgen05": capi.BACKEND_TCGEN05}

TAP_SHAPES = {
    # name -> (channels, frames-per-unit-frame multiplier, extra frames)
    "m_p": (192, 1, 0), "logs_p": (192, 1, 0), "z_p": (192, 1, 0),
    "flow_6": (192, 1, 0), "flow_4": (192, 1, 0), "flow_2": (192, 1, 0), "flow_0": (192, 1, 0),
    "conv_pre": (512, 1, 0), "ups_0": (256, 5, 0), "mrf_0": (256, 5, 0), "ups_1": (128, 20, 0),
    "mrf_1": (128, 20, 0), "conv_post": (72, 20, 1), "y_mb": (4, 80, 0),
}
_TAP_FIELDS = {"m_p": "m_p", "logs_p": "logs_p", "z_p": "z_p", "conv_pre": "conv_pre", "ups_0": "ups0",
               "mrf_0": "mrf0", "ups_1": "ups1", "mrf_1": "mrf1", "conv_post": "conv_post", "y_mb": "y_mb"}
_FLOW_TAPS = ("flow_6", "flow_4", "flow_2", "flow_0")
DECODER_TAPS = ("conv_pre", "ups_0", "mrf_0", "ups_1", "mrf_1", "conv_post", "y_mb")


class InferEngine:
    def __init__(self, module: torch.nn.Module, precision: str, backend: Optional[str], chunk_utts: int) -> None:
        # not an nn.Module attribute: keep the module out of our own __dict__ cycle-free via object.__setattr__
        object.__setattr__(self, "_module", module)
        self._folded = None                                       # (device block, host tail coefficients) behind _model
        self._model: Optional[capi.Model] = None
        self._device: Optional[torch.device] = None
        self._ws: Dict[Tuple[torch.device, int], Tensor] = {}     # one workspace per (device, stream)
        self._watch: list = []                                    # tensors the folded copies were made from
        self._versions = -1
        self.chunk_utts = int(chunk_utts)
        self.configure(precision, backend)

    # ------------------------------------------------------------------ configuration
    def configure(self, precision: str, backend: Optional[str]) -> None:
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}, got {precision!r}")
        if backend is None:
            backend = "fma" if precision == "fp32" else "tcgen05"
        if backend not in _BACKENDS:
            raise ValueError(f"backend must be one of {sorted(_BACKENDS)}, got {backend!r}")
        if backend == "tcgen05" and precision == "fp32":
            raise ValueError("the tcgen05 back end computes with TF32, FP16 or BF16 operands; use precision='tf32' "
                             "(fp32 storage and accumulation) or backend='fma' for exact fp32")
        self.precision, self.backend = precision, backend
        self.invalidate()

    def invalidate(self) -> None:
        self._folded = None
        self._model = None
        self._watch = []
        self._versions = -1

    def _param_versions(self) -> int:
        # in-place edits (p.data.copy_, optimiser steps, manual remove-weight-norm) bump a tensor's version
        # counter; their sum is a cheap fingerprint of "the weights the fold was made from are still the weights"
        return sum(t._version for t in self._watch)

    # ------------------------------------------------------------------ folded weights
    def _ensure_model(self, device: torch.device) -> capi.Model:
        if device.type != "cuda":
            raise capi.QvcError("SynthesizerTrn.infer runs on a B200 only: move the module and its inputs to "
                                "a CUDA device (there is no CPU path)")
        lib = capi.load()
        if self._model is not None and self._device == device and self._param_versions() != self._versions:
            self.invalidate()                      # a parameter was edited in place since the fold
        if self._model is None or self._device != device:
            capi.check(lib.qvc_check_device(device.index if device.index is not None else torch.cuda.current_device()),
                       "qvc_check_device")
            sd = {k: v for k, v in self._module.state_dict().items() if not k.startswith("enc_q.")}
            for k, v in sd.items():
                if v.device != device:
                    raise capi.QvcError(f"parameter {k} lives on {v.device}, inputs on {device}")
            opf = _PRECISIONS[self.precision]
            # The fold is host arithmetic (fp64, once per load) inside the library (qvc_prepare_weights, csrc/fold.cu): the
            # device only ever runs this library's own kernels.  The folded block is uploaded on the current stream, which
            # the call synchronises, so any other stream may use the model struct afterwards without an ordering of its own.
            self._watch = list(sd.values())
            self._versions = self._param_versions()
            entries, keep = capi.state_entries(sd)
            nbytes = int(lib.qvc_prepared_bytes(opf))
            block = torch.empty(nbytes, dtype=torch.uint8, device=device)
            tail_host = torch.zeros(16 + 272, dtype=torch.float32)
            model = capi.Model()
            with torch.cuda.device(device):
                capi.check(lib.qvc_prepare_weights(entries, len(entries), opf, _BACKENDS[self.backend], block.data_ptr(), nbytes,
                                                   tail_host.data_ptr(), C.byref(model), self._stream(device)),
                           "qvc_prepare_weights")
            del keep
            model.chunk_utts = self.chunk_utts
            self._folded = (block, tail_host)            # what the pointers in `model` refer to
            self._model = model
            self._device = device
        return self._model

    def _workspace(self, device: torch.device, nbytes: int) -> Tensor:
        # keyed by the current stream: calls enqueued on different streams run concurrently and must not share
        # scratch memory (include/qvc_b200.h: re-entrant across streams when each stream has its own workspace)
        with torch.cuda.device(device):
            key = (device, torch.cuda.current_stream(device).cuda_stream)
            ws = self._ws.get(key)
            if ws is None or ws.numel() < nbytes:
                self._ws.pop(key, None)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
                self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _f32c(t: Tensor, name: str, device: torch.device) -> Tensor:
        if t.device != device:
            raise capi.QvcError(f"{name} is on {t.device}, expected {device}")
        if t.dtype != torch.float32:
            t = t.float()
        return t.contiguous()

    @staticmethod
    def _stream(device: torch.device) -> int:
        return torch.cuda.current_stream(device).cuda_stream

    @staticmethod
    def _lengths(lengths: Optional[Tensor], B: int, T: int, device: torch.device) -> Optional[Tensor]:
        """(B,) live frames per utterance as a device int32 tensor; range-checked on the host when it is a host
        tensor (a device tensor is trusted: reading it back would synchronise the stream)."""
        if lengths is None:
            return None
        if lengths.dim() != 1 or lengths.shape[0] != B or lengths.dtype.is_floating_point:
            raise ValueError(f"lengths must be an integer tensor of shape ({B},), got {lengths.dtype} {tuple(lengths.shape)}")
        if lengths.device.type == "cpu" and B and (int(lengths.min()) < 1 or int(lengths.max()) > T):
            raise ValueError(f"lengths must lie in [1, {T}], got [{int(lengths.min())}, {int(lengths.max())}]")
        return lengths.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()

    def _make_taps(self, taps: Optional[Dict[str, Tensor]], names, B: int, T: int, n_embed: int,
                   device: torch.device) -> Tuple[Optional[capi.Taps], Dict[str, Tensor]]:
        if taps is None:
            return None, {}
        st = capi.Taps()
        out: Dict[str, Tensor] = {}
        for name in names:
            if name == "g":
                out["g"] = torch.empty(n_embed, 256, device=device)
                st.g = out["g"].data_ptr()
                continue
            ch, mul, extra = TAP_SHAPES[name]
            out[name] = torch.empty(B, ch, mul * T + extra, device=device)
            if name in _FLOW_TAPS:
                st.flow[_FLOW_TAPS.index(name)] = out[name].data_ptr()
            else:
                setattr(st, _TAP_FIELDS[name], out[name].data_ptr())
        return st, out

    # ------------------------------------------------------------------ entry points
    def infer(self, unit: Tensor, mel: Optional[Tensor], noise: Optional[Tensor] = None,
              taps: Optional[Dict[str, Tensor]] = None, g: Optional[Tensor] = None,
              lengths: Optional[Tensor] = None) -> Tensor:
        if unit.dim() != 3 or unit.shape[1] != 256:
            raise ValueError(f"unit must be (B, 256, T), got {tuple(unit.shape)}")
        device = unit.device
        model = self._ensure_model(device)
        lib = capi.load()
        B, _, T = unit.shape
        if B == 0 or T == 0:
            return torch.empty(B, 1, 320 * T, device=device)
        unit = self._f32c(unit, "unit", device)
        lens = self._lengths(lengths, B, T, device)
        if noise is None:
            noise = torch.randn((B, 192, T), device=device, dtype=torch.float32)     # models.py:94
        else:
            if tuple(noise.shape) != (B, 192, T):
                raise ValueError(f"noise must be {(B, 192, T)}, got {tuple(noise.shape)}")
            noise = self._f32c(noise, "noise", device)
        if g is not None:
            g = self._f32c(g.reshape(g.shape[0], -1), "g", device)
            if g.shape[1] != 256 or g.shape[0] not in (1, B):
                raise ValueError(f"g must be (1|B, 256), got {tuple(g.shape)}")
            mel_b, mel_t, mel_ptr, g_ptr, n_embed = g.shape[0], 0, None, g.data_ptr(), g.shape[0]
        else:
            if mel is None or mel.dim() != 3 or mel.shape[1] != 80:
                raise ValueError(f"mel must be (Bm, 80, Tm), got {None if mel is None else tuple(mel.shape)}")
            mel = self._f32c(mel, "mel", device)
            mel_b, mel_t = mel.shape[0], mel.shape[2]
            if mel_t > 128 and mel_b != 1:
                # same failure class as the reference: embed_utterance stacks (W, Bm, 128, 80).squeeze(1)
                # and nn.LSTM rejects the 4-D input (models.py:536)
                raise ValueError(f"mel longer than 128 frames must have batch 1, got {mel_b}")
            n_embed = 1 if mel_t > 128 else mel_b
            if n_embed not in (1, B):
                raise ValueError(f"{n_embed} speaker embeddings cannot broadcast over {B} utterances")
            mel_ptr, g_ptr = mel.data_ptr(), None
        wave = torch.empty(B, 1, 320 * T, device=device, dtype=torch.float32)
        names = ("g",) + tuple(TAP_SHAPES) if taps is not None else ()
        tap_struct, tap_out = self._make_taps(taps, names, B, T, n_embed, device)
        need = lib.qvc_infer_workspace_bytes(C.byref(model), B, T, mel_b, mel_t)
        ws = self._workspace(device, need)
        with torch.cuda.device(device):
            status = lib.qvc_infer(C.byref(model), unit.data_ptr(), mel_ptr, noise.data_ptr(), g_ptr,
                                   lens.data_ptr() if lens is not None else None, B, T, mel_b, mel_t, wave.data_ptr(), C.byref(tap_struct) if tap_struct is not None else None,
                                   ws.data_ptr(), ws.numel(), self._stream(device))
        capi.check(status, "qvc_infer")
        if taps is not None:
            if "g" in tap_out:
                tap_out["g"] = tap_out["g"].unsqueeze(-1)
            taps.update(tap_out)
            taps["wave"] = wave
        return wave

    def embed(self, mel: Tensor) -> Tensor:
        if mel.dim() != 3 or mel.shape[1] != 80:
            raise ValueError(f"mel must be (Bm, 80, Tm), got {tuple(mel.shape)}")
        device = mel.device
        model = self._ensure_model(device)
        lib = capi.load()
        mel = self._f32c(mel, "mel", device)
        bm, _, tm = mel.shape
        if tm > 128 and bm != 1:
            raise ValueError(f"mel longer than 128 frames must have batch 1, got {bm}")
        n_embed = 1 if tm > 128 else bm
        g = torch.empty(n_embed, 256, device=device)
        ws = self._workspace(device, int(lib.qvc_spk_workspace_bytes(bm, tm)) + 256)
        base = (ws.data_ptr() + 255) & ~255
        with torch.cuda.device(device):
            status = lib.qvc_spk_embed(C.byref(model.spk), mel.data_ptr(), bm, tm, g.data_ptr(), base,
                                       ws.numel() - (base - ws.data_ptr()), self._stream(device))
        capi.check(status, "qvc_spk_embed")
        return g

    def decode(self, z: Tensor, g: Tensor, taps: Optional[Dict[str, Tensor]] = None,
               lengths: Optional[Tensor] = None) -> Tensor:
        if z.dim() != 3 or z.shape[1] != 192:
            raise ValueError(f"z must be (B, 192, T), got {tuple(z.shape)}")
        device = z.device
        model = self._ensure_model(device)
        lib = capi.load()
        B, _, T = z.shape
        z = self._f32c(z, "z", device)
        lens = self._lengths(lengths, B, T, device)
        g = self._f32c(g.reshape(g.shape[0], -1), "g", device)
        if g.shape[1] != 256 or g.shape[0] not in (1, B):
            raise ValueError(f"g must be (1|B, 256[, 1]), got {tuple(g.shape)}")
        wave = torch.empty(B, 1, 320 * T, device=device, dtype=torch.float32)
        tap_struct, tap_out = self._make_taps(taps, DECODER_TAPS if taps is not None else (), B, T, g.shape[0], device)
        need = lib.qvc_infer_workspace_bytes(C.byref(model), B, T, 0, 0)
        ws = self._workspace(device, need)
        with torch.cuda.device(device):
            status = lib.qvc_decode(C.byref(model), z.data_ptr(), g.data_ptr(), g.shape[0],
                                    lens.data_ptr() if lens is not None else None, B, T, wave.data_ptr(),
                                    C.byref(tap_struct) if tap_struct is not None else None, ws.data_ptr(),
                                    ws.numel(), self._stream(device))
        capi.check(status, "qvc_decode")
        if taps is not None:
            taps.update(tap_out)
            taps["wave"] = wave
        return wave
