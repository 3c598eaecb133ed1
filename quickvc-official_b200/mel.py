"""Target-speaker mel front end with the reference's signature: `wave_to_mel` (mel_processing.py:79-98).

`convert.py:75-77` calls `wave_to_mel(wav_tgt, 1280, 80, 16000, 320, 1280, 0.0, None)` to produce the `mel` argument of
`SynthesizerTrn.infer`.  The reference builds its filterbank with librosa (not a dependency here) and its spectrum with
`torch.stft`; this module computes both constant operands once per (device, parameters) -- the Hann-windowed DFT basis
and the Slaney mel filterbank, restated from librosa.filters.mel(htk=False, norm='slaney') -- and runs the waveform
through `qvc_wave_to_mel` (include/qvc_b200.h): reflect padding, STFT as an exact-fp32 GEMM over overlapping rows of the
padded waveform, magnitude, filterbank and log in CUDA.  No CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import capi

_cache: Dict[Tuple, Tuple[capi.MelWeights, Tensor, Tensor]] = {}
_ws: Dict[Tuple[torch.device, int], Tensor] = {}     # one workspace per (device, stream)


def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax: Optional[float] = None) -> Tensor:
    """(n_mels, 1 + n_fft // 2) float64, as librosa.filters.mel(sr=sr, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax):
    Slaney's auditory-toolbox mel scale (linear below 1 kHz, logarithmic above) with area ("slaney") normalisation."""
    fmax = sr / 2.0 if fmax is None else float(fmax)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def hz_to_mel(f: float) -> float:
        return f / f_sp if f < min_log_hz else min_log_mel + math.log(f / min_log_hz) / logstep

    def mel_to_hz(m: Tensor) -> Tensor:
        lin = m * f_sp
        return torch.where(m >= min_log_mel, min_log_hz * torch.exp(logstep * (m - min_log_mel)), lin)

    mels = torch.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2, dtype=torch.float64)
    mel_f = mel_to_hz(mels)
    fftfreqs = torch.linspace(0.0, sr / 2.0, 1 + n_fft // 2, dtype=torch.float64)
    fdiff = mel_f[1:] - mel_f[:-1]
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = torch.clamp(torch.minimum(lower, upper), min=0.0)
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return weights * enorm[:, None]


def windowed_dft_basis(n_fft: int, win_size: int) -> Tensor:
    """(round_up(2 bins, 16), n_fft) float64: rows [0, bins) = w[n] cos(2 pi k n / N), rows [bins, 2 bins) = -w[n] sin(...);
    w = periodic Hann of win_size, zero-padded and centred in n_fft as torch.stft does."""
    bins = n_fft // 2 + 1
    w = torch.hann_window(win_size, dtype=torch.float64)
    if win_size < n_fft:
        left = (n_fft - win_size) // 2
        w = torch.nn.functional.pad(w, (left, n_fft - win_size - left))
    n = torch.arange(n_fft, dtype=torch.float64)
    k = torch.arange(bins, dtype=torch.float64)
    ang = 2.0 * math.pi * torch.outer(k, n) / n_fft
    rows16 = (2 * bins + 15) // 16 * 16
    basis = torch.zeros(rows16, n_fft, dtype=torch.float64)
    basis[:bins] = torch.cos(ang) * w
    basis[bins:2 * bins] = -torch.sin(ang) * w
    return basis


def _weights(device: torch.device, n_fft: int, num_mels: int, sr: int, hop: int, win: int, fmin, fmax):
    key = (device, n_fft, num_mels, sr, hop, win, float(fmin or 0.0), fmax)
    if key not in _cache:
        if win > n_fft:
            raise ValueError(f"win_size {win} > n_fft {n_fft}")
        basis = windowed_dft_basis(n_fft, win).to(torch.float32).to(device).contiguous()
        fb_t = slaney_mel_filterbank(sr, n_fft, num_mels, float(fmin or 0.0), fmax).t().to(torch.float32).to(device).contiguous()
        w = capi.MelWeights(basis.data_ptr(), fb_t.data_ptr(), n_fft, hop, num_mels, 0)
        _cache[key] = (w, basis, fb_t)
    return _cache[key][0]


@torch.no_grad()
def wave_to_mel(y: Tensor, n_fft: int, num_mels: int, sampling_rate: int, hop_size: int, win_size: int, fmin, fmax,
                center: bool = False) -> Tensor:
    """y (B, T) fp32 on a CUDA device -> (B, num_mels, frames) log-mel, as the reference's function of this name."""
    if center:
        raise NotImplementedError("the reference calls wave_to_mel with center=False (convert.py:75-77)")
    if y.dim() != 2:
        raise ValueError(f"y must be (B, T), got {tuple(y.shape)}")
    device = y.device
    if device.type != "cuda":
        raise capi.QvcError("wave_to_mel runs on a B200 only: move the waveform to a CUDA device (there is no CPU path)")
    lib = capi.load()
    w = _weights(device, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax)
    y = y.to(torch.float32).contiguous()
    B, T = y.shape
    pad = (n_fft - hop_size) // 2
    if T <= pad:
        raise RuntimeError(f"waveform of {T} samples is not longer than the reflect padding {pad}")   # torch's pad raises too
    frames = lib.qvc_mel_frames(C.byref(w), T)
    mel = torch.empty(B, num_mels, frames, device=device, dtype=torch.float32)
    need = lib.qvc_mel_workspace_bytes(C.byref(w), B, T)
    with torch.cuda.device(device):
        key = (device, torch.cuda.current_stream(device).cuda_stream)
        ws = _ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=device)
            _ws[key] = ws
        st = lib.qvc_wave_to_mel(C.byref(w), y.data_ptr(), B, T, mel.data_ptr(), ws.data_ptr(), ws.numel(),
                                 torch.cuda.current_stream(device).cuda_stream)
    capi.check(st, "qvc_wave_to_mel")
    return mel
