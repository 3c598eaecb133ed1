"""One small conversion per precision + one fused WN layer + one pair-kernel convolution: for compute-sanitizer."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import synth  # noqa: E402
from gpu_util import make_args, op_dtype, stream, to_op  # noqa: E402
from quickvc_official_b200 import SynthesizerTrn, capi  # noqa: E402

os.environ["QVC_TC_2CTA_FORCE"] = "1"          # exercise the CTA-pair kernels on small shapes too
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "quickvc_model_config.json")))
shapes = {k: tuple(v) for k, v in json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json"))).items()}
sd = synth.synthetic_state_dict(shapes, 0)
dev = "cuda:0"
for precision in ("tf32", "bf16"):
    net = SynthesizerTrn(641, 32, **cfg, precision=precision).eval()
    net.load_state_dict(sd)
    net = net.to(dev)
    for B, T in ((2, 150), (1, 37)):
        unit, mel, noise = synth.synthetic_inputs(B, T, 1, 200, 0)
        w = net.infer(unit.to(dev), mel.to(dev), noise=noise.to(dev))
        torch.cuda.synchronize()
        print(precision, B, T, tuple(w.shape), float(w.abs().max()))
print("launches", capi.launch_count())
