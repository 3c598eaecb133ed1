"""CPU oracle for the target-mel front end `wave_to_mel` -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates /root/reference/mel_processing.py:15-98 with the same torch calls the reference makes (reflect pad,
torch.stft(center=False), sqrt(re^2 + im^2 + 1e-6), filterbank matmul, log(clamp 1e-5)).  The one thing the reference
takes from a dependency that is absent here is the filterbank, `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`
(pyproject.toml pins librosa only loosely; mel_processing.py:4,66).  Its published algorithm (Slaney auditory-toolbox
mel scale, area normalisation) is restated below in numpy-style float64, independently of the product's torch
implementation, and pinned in tests/test_oracle.py against torchaudio.functional.melscale_fbanks(norm="slaney",
mel_scale="slaney"), which torchaudio documents as the librosa-compatible filterbank.  The reference module itself cannot
be imported here (librosa missing), so parity of this row is pinned to torch.stft + that filterbank, not to a run of the
reference: "parity partially pinned" (DESIGN.md).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def librosa_mel(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney') -> (n_mels, 1 + n_fft//2) float32."""
    fmax = sr / 2.0 if fmax is None else float(fmax)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - 0.0) / f_sp
    logstep = np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asanyarray(f, dtype=np.float64)
        mel = (f - 0.0) / f_sp
        log_t = f >= min_log_hz
        out = mel.copy()
        out[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
        return out

    def mel_to_hz(m):
        m = np.asanyarray(m, dtype=np.float64)
        f = f_sp * m
        log_t = m >= min_log_mel
        f[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
        return f

    lo, hi = hz_to_mel(np.array([fmin, fmax]))
    mel_f = mel_to_hz(np.linspace(lo, hi, n_mels + 2))
    fftfreqs = np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


def wave_to_mel(y: torch.Tensor, n_fft: int, num_mels: int, sampling_rate: int, hop_size: int, win_size: int, fmin, fmax,
                center: bool = False, dtype=torch.float32) -> torch.Tensor:
    """mel_processing.py:15-98 on the CPU; `dtype=torch.float64` gives the high-precision reference of the same recipe."""
    y = y.to(dtype)
    window = torch.hann_window(win_size).to(dtype)                                   # mel_processing.py:34
    pad = int((n_fft - hop_size) / 2)
    y = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)   # :46-47
    spec = torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=window, center=center,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)   # :50-51
    spec = torch.sqrt(spec.real.pow(2) + spec.imag.pow(2) + 1e-6)                    # :54
    mel_basis = torch.from_numpy(librosa_mel(sampling_rate, n_fft, num_mels, fmin or 0.0, fmax)).to(dtype)   # :66-67
    return torch.log(torch.clamp(torch.matmul(mel_basis, spec), min=1e-5))           # :70-71, :8
