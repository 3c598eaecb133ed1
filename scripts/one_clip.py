"""One 5 s clip through infer (B=1, T=250), a few times: the command the latency launch list is taken from."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from quickvc_official_b200 import SynthesizerTrn  # noqa: E402

cfg = bench.model_cfg()
dev = torch.device("cuda:0")
net = SynthesizerTrn(641, 32, precision=os.environ.get("QVC_PRECISION", "tf32"), **cfg).eval()
net.load_state_dict(bench.random_init_state_dict(cfg))
net = net.to(dev)
g = torch.Generator().manual_seed(1)
T = int(os.environ.get("CLIP_T", "250"))
unit = torch.randn(1, 256, T, generator=g).to(dev)
mel = (torch.randn(1, 80, 250, generator=g) * 2 - 5).to(dev)
for _ in range(int(os.environ.get("CLIP_N", "3"))):
    w = net.infer(unit, mel)
torch.cuda.synchronize()
print("ok", tuple(w.shape), float(w.abs().max()))
