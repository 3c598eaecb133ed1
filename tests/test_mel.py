"""Target-mel front end: oracle pinning on the CPU, CUDA parity on a B200."""
import math

import numpy as np
import pytest
import torch

from oracle import mel_oracle
from quickvc_official_b200 import mel as qmel

ARGS = (1280, 80, 16000, 320, 1280, 0.0, None)           # convert.py:75-77 with configs/quickvc.json:24-32


def test_oracle_filterbank_matches_torchaudio_slaney():
    """librosa.filters.mel is not installed; torchaudio documents melscale_fbanks(norm='slaney', mel_scale='slaney')
    as the librosa-compatible filterbank -- pin the oracle's restatement and the product's to it."""
    ta = pytest.importorskip("torchaudio")
    want = ta.functional.melscale_fbanks(n_freqs=641, f_min=0.0, f_max=8000.0, n_mels=80, sample_rate=16000,
                                         norm="slaney", mel_scale="slaney").t().double()
    got_oracle = torch.from_numpy(mel_oracle.librosa_mel(16000, 1280, 80, 0.0, None)).double()
    got_product = qmel.slaney_mel_filterbank(16000, 1280, 80, 0.0, None)
    assert got_oracle.shape == (80, 641)
    assert float((got_oracle - want).abs().max()) < 1e-7
    assert float((got_product - want).abs().max()) < 1e-7
    # known properties of the Slaney bank: triangles peak in order, area normalisation (each filter integrates to ~1 Hz^-1 * 2)
    peaks = got_oracle.argmax(dim=1)
    assert bool((peaks[1:] >= peaks[:-1]).all())


def test_oracle_shapes_and_fp64_agreement():
    g = torch.Generator().manual_seed(0)
    y = torch.rand(2, 16000, generator=g) * 1.6 - 0.8
    m32 = mel_oracle.wave_to_mel(y, *ARGS)
    m64 = mel_oracle.wave_to_mel(y, *ARGS, dtype=torch.float64)
    assert m32.shape == (2, 80, 50)                       # 16000 / 320 frames, as in the reference
    assert float((m32.double() - m64).abs().max()) < 2e-4
    with pytest.raises(RuntimeError):                     # reflect padding needs more than 480 samples
        mel_oracle.wave_to_mel(torch.zeros(1, 480), *ARGS)


def test_windowed_dft_basis_is_the_stft():
    g = torch.Generator().manual_seed(1)
    frame = torch.randn(1280, generator=g, dtype=torch.float64)
    basis = qmel.windowed_dft_basis(1280, 1280)
    want = torch.fft.rfft(frame * torch.hann_window(1280, dtype=torch.float64))
    got = basis @ frame
    assert float((got[:641] - want.real).abs().max()) < 1e-9 and float((got[641:1282] - want.imag).abs().max()) < 1e-9
    assert float(basis[1282:].abs().max()) == 0.0


def test_wave_to_mel_has_no_cpu_path():
    from quickvc_official_b200 import capi
    with pytest.raises(capi.QvcError):
        qmel.wave_to_mel(torch.zeros(1, 4000), *ARGS)


@pytest.mark.gpu
@pytest.mark.parametrize("batch,samples", [(1, 160000), (1, 80000), (3, 16000), (2, 12345), (1, 481), (1, 1280)])
def test_wave_to_mel_matches_oracle_on_b200(batch, samples):
    g = torch.Generator().manual_seed(samples)
    y = torch.rand(batch, samples, generator=g) * 1.9 - 0.95
    y[:, : samples // 3] *= 0.01                           # a quiet stretch: exercises the 1e-6 / 1e-5 floors
    want = mel_oracle.wave_to_mel(y, *ARGS, dtype=torch.float64)
    got = qmel.wave_to_mel(y.to("cuda:0"), *ARGS)
    torch.cuda.synchronize()
    assert got.shape == want.shape and got.dtype == torch.float32
    err = float((got.cpu().double() - want).abs().max())
    assert err < 2e-4, err                                 # log-mel, fp32 GEMM of 1280 terms vs fp64
    ref32 = mel_oracle.wave_to_mel(y, *ARGS)               # the reference's own fp32 arithmetic is no closer to fp64
    assert err < 4 * float((ref32.double() - want).abs().max()) + 1e-5


@pytest.mark.gpu
def test_mel_feeds_infer_like_convert_py(sd, model_cfg):
    """convert.py:75-81: mel_tgt = wave_to_mel(wav_tgt, ...); audio = net_g.infer(unit, mel_tgt)."""
    import synth
    from quickvc_official_b200 import SynthesizerTrn
    from oracle import qvc_oracle
    g = torch.Generator().manual_seed(5)
    wav_tgt = (torch.rand(1, 48000, generator=g) * 2 - 1) * 0.3
    unit, _, noise = synth.synthetic_inputs(1, 40, 1, 150, 0)
    net = SynthesizerTrn(641, 32, **model_cfg).eval()
    net.load_state_dict(sd)
    net = net.to("cuda:0")
    mel = qmel.wave_to_mel(wav_tgt.to("cuda:0"), *ARGS)
    wave = net.infer(unit.to("cuda:0"), mel, noise=noise.to("cuda:0"))
    ref = qvc_oracle.infer(sd, unit, mel_oracle.wave_to_mel(wav_tgt, *ARGS), noise)
    assert float((wave.cpu() - ref).abs().max()) < 1e-4


# ------------------------------------------------------------------------------------------------
# Pin to the reference itself: tests/golden/mel_*.npz are outputs of the UNMODIFIED /root/reference/mel_processing.py
# (tests/golden/make_mel_golden.py: the module imported as is, only the absent `librosa.filters.mel` stubbed with
# torchaudio's librosa-compatible Slaney bank).
# ------------------------------------------------------------------------------------------------
def _mel_golden(name):
    import os
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    import make_mel_golden
    batch, samples, seed = make_mel_golden.CASES[name]
    with np.load(os.path.join(GOLDEN, f"mel_{name}.npz")) as z:
        return make_mel_golden.waves(batch, samples, seed), torch.from_numpy(z["mel"])


MEL_CASES = ("10s", "1s_b3", "ragged", "min")


@pytest.mark.parametrize("name", MEL_CASES)
def test_oracle_matches_reference_mel_golden(name):
    y, gold = _mel_golden(name)
    got = mel_oracle.wave_to_mel(y, *ARGS)
    assert got.shape == gold.shape
    # same torch calls on the same machine class; the only independent piece is the restated filterbank
    assert float((got - gold).abs().max()) < 2e-5, float((got - gold).abs().max())


def test_oracle_matches_live_reference_mel():
    import os
    import sys
    from conftest import GOLDEN, REFERENCE
    if not os.path.isdir(REFERENCE):
        pytest.skip("reference tree not mounted")
    pytest.importorskip("torchaudio")
    sys.path.insert(0, GOLDEN)
    import make_mel_golden
    make_mel_golden.stub_librosa()
    sys.path.insert(0, REFERENCE)
    import mel_processing
    y = make_mel_golden.waves(2, 20000, 11)
    want = mel_processing.wave_to_mel(y, *ARGS)
    got = mel_oracle.wave_to_mel(y, *ARGS)
    assert float((got - want).abs().max()) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name", MEL_CASES)
def test_wave_to_mel_matches_reference_golden_on_b200(name):
    y, gold = _mel_golden(name)
    got = qmel.wave_to_mel(y.to("cuda:0"), *ARGS)
    torch.cuda.synchronize()
    assert got.shape == gold.shape and got.dtype == torch.float32
    # fp32 GEMM of 1280 terms in a different summation order than torch.stft's FFT, then a log
    assert float((got.cpu() - gold).abs().max()) < 3e-4, float((got.cpu() - gold).abs().max())
