"""Generate tests/golden/mel_*.npz by running the UNMODIFIED reference `mel_processing.wave_to_mel`
(/root/reference/mel_processing.py:79-98) in-process.

Run in the authoring container only (the reference does not travel to the GPU box):
    python tests/golden/make_mel_golden.py
The reference module imports `librosa.filters.mel` at module level (mel_processing.py:4) and librosa is not installed
here.  A stub module `librosa.filters` is therefore injected into sys.modules whose `mel(sr, n_fft, n_mels, fmin, fmax)`
returns torchaudio's `melscale_fbanks(norm="slaney", mel_scale="slaney")` -- the filterbank torchaudio documents as
librosa-compatible (htk=False, norm="slaney" are librosa's defaults) -- transposed to librosa's (n_mels, 1 + n_fft//2)
float32 layout.  Everything else (reflect pad, torch.stft, magnitude, matmul, log-clamp) is the reference's own code.
"""
import os
import sys
import types

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QVC_REFERENCE", "/root/reference")
ARGS = (1280, 80, 16000, 320, 1280, 0.0, None)           # convert.py:75-77 with configs/quickvc.json:24-32

CASES = {               # name: (batch, samples, seed)
    "10s": (1, 160000, 0),
    "1s_b3": (3, 16000, 1),
    "ragged": (2, 12345, 2),
    "min": (1, 481, 3),
}


def stub_librosa():
    def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw):
        assert not kw, f"stub only covers librosa's defaults, got {kw}"
        fb = torchaudio.functional.melscale_fbanks(n_freqs=1 + n_fft // 2, f_min=float(fmin),
                                                   f_max=float(sr) / 2 if fmax is None else float(fmax), n_mels=n_mels,
                                                   sample_rate=sr, norm="slaney", mel_scale="slaney")
        return fb.t().contiguous().numpy().astype(np.float32)
    lib = types.ModuleType("librosa")
    filt = types.ModuleType("librosa.filters")
    filt.mel = mel
    lib.filters = filt
    sys.modules["librosa"] = lib
    sys.modules["librosa.filters"] = filt


def waves(batch, samples, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, samples, generator=g) * 1.6 - 0.8


def main():
    stub_librosa()
    sys.path.insert(0, REF)
    import mel_processing                                   # the reference's module, unmodified
    for name, (batch, samples, seed) in CASES.items():
        y = waves(batch, samples, seed)
        m = mel_processing.wave_to_mel(y, *ARGS)
        np.savez_compressed(os.path.join(HERE, f"mel_{name}.npz"), mel=m.numpy())
        print(name, tuple(m.shape), float(m.min()), float(m.max()))


if __name__ == "__main__":
    main()
