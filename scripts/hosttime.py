import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, bench
from quickvc_official_b200 import SynthesizerTrn
cfg = bench.model_cfg(); sd = bench.random_init_state_dict(cfg); dev = torch.device("cuda:0")
net = SynthesizerTrn(641, 32, **cfg).eval(); net.load_state_dict(sd); net = net.to(dev)
for T in (25, 250):
    g = torch.Generator().manual_seed(1)
    unit = torch.randn(1, 256, T, generator=g).to(dev); mel = (torch.randn(1, 80, 250, generator=g) * 2 - 5).to(dev)
    noise = torch.randn(1, 192, T, generator=g).to(dev); emb = net.embed_speaker(mel)
    for _ in range(10): net.infer_with_embedding(unit, emb, noise=noise)
    torch.cuda.synchronize()
    ts = []
    for _ in range(50):
        torch.cuda.synchronize(); t0 = time.perf_counter(); net.infer_with_embedding(unit, emb, noise=noise); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t0))
    ts.sort(); print(T, "host enqueue ms", round(ts[25][0] * 1e3, 3), "total ms", round(sorted(t[1] for t in ts)[25] * 1e3, 3))
