"""Batch conversion driver on a B200 (reference convert.py:47-86): every file it writes equals the single-utterance
chain -- trim -> mel -> infer -- run one clip at a time, and the CPU oracle's result for the same clip."""
import os

import numpy as np
import pytest
import torch
from scipy.io import wavfile

from oracle import mel_oracle, qvc_oracle
from quickvc_official_b200 import SynthesizerTrn
from quickvc_official_b200 import convert as cv
from quickvc_official_b200.mel import wave_to_mel

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DATA = dict(sampling_rate=16000, filter_length=1280, hop_length=320, win_length=1280, n_mel_channels=80,
            mel_fmin=0.0, mel_fmax=None)


def _speechlike(rng, seconds, sr):
    n = int(seconds * sr)
    t = np.arange(n) / sr
    env = np.clip(np.sin(2 * np.pi * 1.3 * t + rng.uniform(0, 3)), 0, None) ** 2
    y = env * (0.3 * np.sin(2 * np.pi * 180 * t) + 0.1 * rng.standard_normal(n))
    y[: int(0.2 * sr)] *= 1e-3                        # quiet head and tail for the trim
    y[-int(0.3 * sr):] *= 1e-3
    return y.astype(np.float32)


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("convert")
    rng = np.random.default_rng(5)
    (d / "units").mkdir()
    (d / "wav").mkdir()
    wavfile.write(d / "wav" / "tgtA.wav", 16000, (_speechlike(rng, 3.1, 16000) * 32767).astype(np.int16))
    wavfile.write(d / "wav" / "tgtB.wav", 22050, _speechlike(rng, 2.4, 22050))          # resampled on load
    frames = {"s0": 61, "s1": 25, "s2": 61, "s3": 140, "s4": 61, "s5": 33}
    for name, t in frames.items():
        torch.save(torch.from_numpy(rng.standard_normal((1, t, 256)).astype(np.float32)), d / "units" / f"{name}.pt")
    lines = ["a0|wav/s0.wav|{d}/wav/tgtA.wav", "a1|wav/s1.wav|{d}/wav/tgtA.wav", "a2|wav/s2.wav|{d}/wav/tgtA.wav",
             "b3|wav/s3.wav|{d}/wav/tgtB.wav", "a4|wav/s4.wav|{d}/wav/tgtA.wav", "b5|wav/s5.wav|{d}/wav/tgtB.wav",
             "b0|wav/s0.wav|{d}/wav/tgtB.wav"]
    (d / "convert.txt").write_text("\n".join(l.format(d=d) for l in lines) + "\n")
    return d


@pytest.mark.parametrize("streams,ragged", [(1, False), (3, False), (2, True)])
def test_convert_list_matches_single_clip_chain(workdir, sd, model_cfg, streams, ragged):
    hps = cv.HParams(data=DATA, train={"segment_size": 10240}, model=model_cfg)
    net = cv.build_net(hps, None, torch.device(DEV))
    net.load_state_dict(sd)
    conv = cv.Converter(net, hps, units_dir=str(workdir / "units"), streams=streams, max_batch=2 if not ragged else 4,
                        ragged=ragged, max_padding=0.5)
    out = workdir / f"out{streams}{ragged}"
    written = conv.convert_list(str(workdir / "convert.txt"), str(out), noise_seed=100)
    items = cv.read_list(str(workdir / "convert.txt"))
    assert sorted(os.path.basename(p) for p in written) == sorted(t + ".wav" for t, _, _ in items)
    audio_s = sum(torch.load(workdir / "units" / (os.path.basename(s)[:-4] + ".pt")).shape[1] for _, s, _ in items) / 50.0
    assert conv.stats["utterances"] == 7 and conv.stats["targets"] == 2 and conv.stats["audio_seconds"] == pytest.approx(audio_s)
    if ragged:
        # sorted lengths 25 33 61 61 61 61 140, at most 4 per call and 50 % padding: [25 33 61 61] [61 61 140]
        assert conv.stats["calls"] == 2 and conv.stats["padded_audio_seconds"] == pytest.approx((4 * 61 + 3 * 140) / 50.0)
    else:
        # a0, a2, a4 share (target A, 61 frames): batches of 2 + 1; everything else runs alone
        assert conv.stats["calls"] == 6 and conv.stats["padded_audio_seconds"] == pytest.approx(audio_s)

    checked_oracle = False
    for index, (title, src, tgt) in enumerate(items):
        sr, got = wavfile.read(out / f"{title}.wav")
        assert sr == 16000 and got.dtype == np.float32
        # the reference's per-utterance chain (convert.py:62-81), one clip at a time through the public API
        wav_t = torch.from_numpy(cv.trim_silence(cv.load_wave(tgt, 16000), top_db=20)).unsqueeze(0)
        mel = wave_to_mel(wav_t.to(DEV), 1280, 80, 16000, 320, 1280, 0.0, None)
        unit = torch.load(workdir / "units" / (os.path.basename(src)[:-4] + ".pt")).transpose(2, 1).contiguous().to(DEV)
        gen = torch.Generator(device=DEV)
        gen.manual_seed(100 + index)
        noise = torch.randn((1, 192, unit.shape[2]), device=DEV, generator=gen)
        want = net.infer(unit, mel, noise=noise)[0, 0].cpu().numpy()
        assert got.shape == want.shape == (320 * unit.shape[2],)
        assert np.array_equal(got, want), f"{title}: driver output differs from the single-clip call by {np.abs(got - want).max()}"   # bit for bit, ragged batches included
        if not checked_oracle and title == "b5":
            mel_ref = mel_oracle.wave_to_mel(wav_t.double(), 1280, 80, 16000, 320, 1280, 0.0, None)
            ref = qvc_oracle.infer(sd, unit.cpu(), mel_ref.float(), noise.cpu(), dtype=torch.float64)[0, 0].numpy()
            assert np.abs(got - ref).max() <= 1e-4      # fp32-mode tolerance of the path (BASELINE.json north_star)
            checked_oracle = True
    assert checked_oracle


def test_converter_needs_units_or_encoder(sd, model_cfg):
    hps = cv.HParams(data=DATA, train={"segment_size": 10240}, model=model_cfg)
    net = cv.build_net(hps, None, torch.device(DEV))
    with pytest.raises(ValueError):
        cv.Converter(net, hps)
    # a content_encoder callable stands in for hubert_soft.units (convert.py:79)
    calls = []

    def encoder(wav):
        calls.append(tuple(wav.shape))
        return torch.zeros(1, wav.shape[-1] // 320, 256, device=wav.device)

    conv = cv.Converter(net, hps, content_encoder=encoder)
    assert conv.units is None and conv.content_encoder is encoder
