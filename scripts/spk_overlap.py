"""Does the speaker encoder on its side stream slow the prior encoder?  Step time with and without it (B = 64 x 10 s)."""
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
for precision in ("tf32", "bf16"):
    net = SynthesizerTrn(641, 32, **cfg, precision=precision).eval()
    net.load_state_dict(sd)
    net = net.to(dev)
    g = torch.Generator().manual_seed(1)
    unit = torch.randn(64, 256, 500, generator=g).to(dev)
    mel = (torch.randn(1, 80, 500, generator=g) * 2 - 5).to(dev)
    noise = torch.randn(64, 192, 500, generator=g).to(dev)
    emb = net.embed_speaker(mel)
    for name, fn in (("with speaker encoder", lambda: net.infer(unit, mel, noise=noise)),
                     ("cached embedding", lambda: net.infer_with_embedding(unit, emb, noise=noise))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            fn()
        e.record()
        torch.cuda.synchronize()
        print(f"{precision} {name:22s} {s.elapsed_time(e) / 10:.3f} ms per step", flush=True)
