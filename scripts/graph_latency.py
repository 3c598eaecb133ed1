"""CUDA-graph capture of one infer call (fork / join of the speaker-encoder side stream included) and its replay latency."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from quickvc_official_b200 import SynthesizerTrn  # noqa: E402

cfg = bench.model_cfg()
sd = bench.random_init_state_dict(cfg)
dev = torch.device("cuda:0")
net = SynthesizerTrn(641, 32, **cfg).eval()
net.load_state_dict(sd)
net = net.to(dev)
for T in (250, 25):
    g = torch.Generator().manual_seed(1)
    unit = torch.randn(1, 256, T, generator=g).to(dev)
    mel = (torch.randn(1, 80, 250, generator=g) * 2 - 5).to(dev)
    noise = torch.randn(1, 192, T, generator=g).to(dev)
    emb = net.embed_speaker(mel)
    for name, fn in (("infer(unit, mel)", lambda: net.infer(unit, mel, noise=noise)),
                     ("cached speaker", lambda: net.infer_with_embedding(unit, emb, noise=noise))):
        ref = fn().clone()
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = fn()
        graph.replay()
        torch.cuda.synchronize()
        same = bool(torch.equal(out, ref))
        lat = []
        for i in range(210):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            if i >= 10:
                lat.append(a.elapsed_time(b))
        lat.sort()
        print(f"T={T:3d} {name:18s} graph replay p50 {lat[len(lat) // 2]:.3f} ms  p99 {lat[int(len(lat) * 0.99)]:.3f} ms  identical={same}", flush=True)
