"""Ragged conversion through the batch driver: audio-seconds per second over a list of clips of mixed lengths,
by number of CUDA streams (every clip keeps the reference's exact single-utterance arithmetic)."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy.io import wavfile  # noqa: E402

import bench  # noqa: E402
from quickvc_official_b200 import convert as cv  # noqa: E402

N = int(os.environ.get("CVB_CLIPS", "256"))
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
cfg = bench.model_cfg()
data = dict(sampling_rate=16000, filter_length=1280, hop_length=320, win_length=1280, n_mel_channels=80, mel_fmin=0.0, mel_fmax=None)
hps = cv.HParams(data=data, train={"segment_size": 10240}, model=cfg)
dev = torch.device("cuda:0")
net = cv.build_net(hps, None, dev, prec)
net.load_state_dict(bench.random_init_state_dict(cfg))
rng = np.random.default_rng(0)
with tempfile.TemporaryDirectory() as d:
    os.makedirs(f"{d}/units")
    for s in range(4):
        wavfile.write(f"{d}/tgt{s}.wav", 16000, (rng.standard_normal(16000 * 5) * 0.1).astype(np.float32))
    lines = []
    total = 0
    for i in range(N):
        t = int(rng.integers(100, 501))                 # 2 .. 10 s
        total += t
        torch.save(torch.from_numpy(rng.standard_normal((1, t, 256)).astype(np.float32)), f"{d}/units/s{i}.pt")
        lines.append(f"c{i}|s{i}.wav|{d}/tgt{i % 4}.wav")
    open(f"{d}/list.txt", "w").write("\n".join(lines) + "\n")
    items = cv.read_list(f"{d}/list.txt")
    for streams, ragged in ((1, False), (4, False), (1, True), (2, True)):
        conv = cv.Converter(net, hps, units_dir=f"{d}/units", streams=streams, ragged=ragged)
        for _ in conv.convert(items[:16]):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = sum(1 for _ in conv.convert(items))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        conv.convert_list(f"{d}/list.txt", f"{d}/out{streams}")
        dt_files = time.perf_counter() - t1
        # units already on the device (what a content encoder running on the GPU hands over): no file reads
        mem_items = [(t, conv.source_units(s).transpose(2, 1).contiguous(), g) for t, s, g in items]
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        n2 = sum(1 for _ in conv.convert(mem_items))
        torch.cuda.synchronize()
        dt_mem = time.perf_counter() - t2
        print(f"{prec} ragged={ragged} streams={streams} calls={conv.stats['calls']} padding={conv.stats['padded_audio_seconds'] / conv.stats['audio_seconds'] - 1:.1%}: {n} clips, {total / 50:.0f} audio-s in {dt * 1e3:.0f} ms = {total / 50 / dt:.0f} audio-s/s "
              f"({dt / n * 1e3:.2f} ms per clip); with WAV writing {total / 50 / dt_files:.0f} audio-s/s; "
              f"units resident on the device {total / 50 / dt_mem:.0f} audio-s/s ({dt_mem * 1e3:.0f} ms)", flush=True)
