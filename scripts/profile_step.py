"""One conversion step for profiling under ncu: `python scripts/profile_step.py [precision] [B] [T] [chunk]`.
Runs `warm` untimed calls then one marked call; prints the library launch count of one call."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from quickvc_official_b200 import SynthesizerTrn, capi  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 500
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
warm = int(os.environ.get("QVC_PROFILE_WARM", "1"))
cfg = bench.model_cfg()
net = SynthesizerTrn(641, 32, **cfg, precision=precision, chunk_utts=chunk).eval()
net.load_state_dict(bench.random_init_state_dict(cfg))
dev = torch.device("cuda:0")
net = net.to(dev)
g = torch.Generator().manual_seed(1)
unit = torch.randn(B, 256, T, generator=g).to(dev)
mel = (torch.randn(1, 80, T, generator=g) * 2 - 5).to(dev)
noise = torch.randn(B, 192, T, generator=g).to(dev)
for _ in range(warm):
    net.infer(unit, mel, noise=noise)
torch.cuda.synchronize()
n0 = capi.launch_count()
torch.cuda.cudart().cudaProfilerStart()
net.infer(unit, mel, noise=noise)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("launches per call:", capi.launch_count() - n0)
