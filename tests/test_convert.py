"""Host logic of the batch conversion driver (quickvc_official_b200.convert; reference convert.py:19-86): CPU only."""
import json
import os

import numpy as np
import pytest
import torch
from scipy.io import wavfile

from quickvc_official_b200 import convert as cv


def test_read_list(tmp_path):
    p = tmp_path / "convert.txt"
    p.write_text("title1|a/src1.wav|b/tgt.wav\n\ntitle2|a/src2.wav|b/tgt.wav\n")
    assert cv.read_list(str(p)) == [("title1", "a/src1.wav", "b/tgt.wav"), ("title2", "a/src2.wav", "b/tgt.wav")]
    p.write_text("only|two\n")
    with pytest.raises(ValueError):
        cv.read_list(str(p))


def test_load_wave_formats(tmp_path):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(4000) * 0.2).astype(np.float32)
    wavfile.write(tmp_path / "f32.wav", 16000, x)
    assert np.array_equal(cv.load_wave(str(tmp_path / "f32.wav"), 16000), x)
    i16 = (x * 32767).astype(np.int16)
    wavfile.write(tmp_path / "i16.wav", 16000, i16)
    assert np.array_equal(cv.load_wave(str(tmp_path / "i16.wav"), 16000), i16.astype(np.float32) / 32768.0)
    st = np.stack([i16, -i16 // 2], axis=1)
    wavfile.write(tmp_path / "st.wav", 16000, st)
    mono = cv.load_wave(str(tmp_path / "st.wav"), 16000)
    assert mono.shape == (4000,) and np.allclose(mono, st.astype(np.float32).mean(axis=1) / 32768.0, atol=1e-7)
    # resampling: 22.05 kHz tone keeps its frequency and duration at 16 kHz
    t = np.arange(22050) / 22050.0
    wavfile.write(tmp_path / "hi.wav", 22050, (0.5 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32))
    y = cv.load_wave(str(tmp_path / "hi.wav"), 16000)
    assert y.dtype == np.float32 and abs(len(y) - 16000) <= 1
    spec = np.abs(np.fft.rfft(y[:16000] if len(y) >= 16000 else np.pad(y, (0, 16000 - len(y)))))
    assert int(spec.argmax()) == 440


def _trim_naive(y, top_db, frame=2048, hop=512):
    """Independent restatement, frame by frame (librosa.effects.trim -> feature.rms -> amplitude_to_db)."""
    x = np.pad(y.astype(np.float64), (frame // 2, frame // 2))
    rms = np.array([np.sqrt(np.mean(x[t * hop:t * hop + frame] ** 2)) for t in range(1 + len(y) // hop)])
    mag = np.abs(rms)
    db = 10 * np.log10(np.maximum(1e-10, mag ** 2)) - 10 * np.log10(max(1e-10, mag.max() ** 2))
    nz = np.flatnonzero(db > -top_db)
    if nz.size == 0:
        return 0, 0
    return int(nz[0]) * hop, min(len(y), (int(nz[-1]) + 1) * hop)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_trim_silence_matches_framewise_definition(seed):
    rng = np.random.default_rng(seed)
    n = 30000 + 777 * seed
    y = (rng.standard_normal(n) * 1e-3).astype(np.float32)
    a, b = 5000 + 300 * seed, 21000 - 100 * seed
    y[a:b] += (rng.standard_normal(b - a) * 0.3 * np.hanning(b - a)).astype(np.float32)
    s, e = _trim_naive(y, 20)
    out = cv.trim_silence(y, top_db=20)
    assert 0 < s < a + 2048 and b - 2048 < e <= n
    assert np.array_equal(out, y[s:e])


def test_trim_silence_edges():
    assert cv.trim_silence(np.zeros(0, np.float32)).shape == (0,)
    z = np.zeros(5000, np.float32)
    assert cv.trim_silence(z).shape == (5000,)          # every frame is 0 dB below the (floored) maximum: nothing trimmed
    y = np.zeros(8192, np.float32)
    y[4096] = 1.0                                        # one click: only the frames that contain it survive
    out = cv.trim_silence(y, top_db=20)
    s, e = _trim_naive(y, 20)
    assert (s, e) == (3584, 5632) and np.array_equal(out, y[s:e])


def test_hparams_and_checkpoint(tmp_path):
    cfg = {"train": {"segment_size": 10240}, "data": {"sampling_rate": 16000, "hop_length": 320, "mel_fmax": None},
           "model": {"inter_channels": 192, "resblock": "1"}}
    (tmp_path / "c.json").write_text(json.dumps(cfg))
    hps = cv.get_hparams_from_file(str(tmp_path / "c.json"))
    assert hps.data.sampling_rate == 16000 and hps.model.as_dict() == cfg["model"] and hps.data.mel_fmax is None
    assert "train" in hps and hps["train"].segment_size == 10240

    net = torch.nn.Linear(3, 2)
    saved = {"weight": torch.full((2, 3), 7.0)}          # bias missing from the file: keeps the module's value (utils.py:167-172)
    torch.save({"model": saved, "iteration": 12, "learning_rate": 1e-4}, tmp_path / "g.pth")
    bias = net.bias.detach().clone()
    assert cv.load_checkpoint(str(tmp_path / "g.pth"), net) == 12
    assert torch.equal(net.weight, saved["weight"]) and torch.equal(net.bias, bias)
    with pytest.raises(FileNotFoundError):
        cv.load_checkpoint(str(tmp_path / "missing.pth"), net)


def test_units_directory(tmp_path):
    torch.save(torch.randn(31, 256), tmp_path / "a.pt")
    np.save(tmp_path / "b.npy", np.random.default_rng(0).standard_normal((1, 17, 256)).astype(np.float32))
    np.save(tmp_path / "bad.npy", np.zeros((5, 100), np.float32))
    ud = cv.UnitsDirectory(str(tmp_path))
    cpu = torch.device("cpu")
    assert tuple(ud.for_source("x/y/a.wav", cpu).shape) == (1, 31, 256)
    assert tuple(ud.for_source("b.flac", cpu).shape) == (1, 17, 256)
    with pytest.raises(ValueError):
        ud.for_source("bad.wav", cpu)
    with pytest.raises(FileNotFoundError):
        ud.for_source("nothing.wav", cpu)


def test_converter_refuses_cpu():
    net = torch.nn.Linear(2, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        cv.Converter(net, cv.HParams(data={"sampling_rate": 16000}), units_dir=".")


def test_shard_items_partitions_the_list():
    items = [(f"t{i}", f"s{i}.wav", "tgt.wav") for i in range(11)]
    for world in (1, 2, 3, 8, 16):
        shares = [cv.shard_items(items, r, world) for r in range(world)]
        assert sorted(sum(shares, [])) == sorted(items)
        assert max(map(len, shares)) - min(map(len, shares)) <= 1
    with pytest.raises(ValueError):
        cv.shard_items(items, 2, 2)
