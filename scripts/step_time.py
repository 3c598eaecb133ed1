"""Whole-step time of one configuration: python scripts/step_time.py [precision] [B] [T] [reps] (env knobs apply)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from quickvc_official_b200 import SynthesizerTrn
precision = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 500
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
cfg = bench.model_cfg()
net = SynthesizerTrn(641, 32, **cfg, precision=precision).eval()
net.load_state_dict(bench.random_init_state_dict(cfg))
dev = torch.device("cuda:0"); net = net.to(dev)
g = torch.Generator().manual_seed(1)
unit = torch.randn(B, 256, T, generator=g).to(dev); mel = (torch.randn(1, 80, T, generator=g) * 2 - 5).to(dev)
noise = torch.randn(B, 192, T, generator=g).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): net.infer(unit, mel, noise=noise)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); w = net.infer(unit, mel, noise=noise); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ts.sort()
knobs = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("QVC_"))
print(f"{precision} B={B} T={T} [{knobs}]: median {ts[len(ts)//2]:.3f} ms  min {ts[0]:.3f}  -> {B*T/50/ts[len(ts)//2]*1e3:.0f} audio-s/s  checksum {float(w.double().abs().sum()):.6f}", flush=True)
