#!/usr/bin/env python
"""How wide is the margin of the fp32-mode tolerance (waveform max-abs 1e-4, per-stage rel-L2 1e-3) over weight sets?

CPU only.  For several seeds of two weight families -- tests/synth.py's (the golden fixtures' family) and the reference
constructor's own initialisation as bench.py draws it -- the launch sequence of the engine is emulated with its operand
rounding (tests/emulate.py: TF32 / fp16 / bf16 operands, fp32 accumulation and streams) and compared with the fp64 oracle
on one 5 s utterance (BASELINE.json configs[0] shape).  The emulation rounds where the kernels round; it does not model
accumulation order.  The GPU suite holds the kernels themselves to the same bounds on two seeds at full size
(tests/test_gpu_fullsize.py); this table is the wider, cheaper sweep behind the claim.

    python tests/tolerance_seeds.py [n_seeds] > profiles/r02_tolerance_seeds.md
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402
from emulate import Emu  # noqa: E402
from oracle import qvc_oracle  # noqa: E402
from quickvc_official_b200 import SynthesizerTrn, capi  # noqa: E402


def ctor_state_dict(cfg, seed):
    torch.manual_seed(seed)
    sd = SynthesizerTrn(641, 32, **cfg).state_dict()
    g = torch.Generator().manual_seed(7 + seed)
    for k in sd:
        if ".post." in k and k.startswith("flow."):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.05
    return sd


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "quickvc_model_config.json")))
    shapes = {k: tuple(v) for k, v in json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json"))).items()}
    modes = (("tf32", capi.OPF_TF32), ("fp16", capi.OPF_F16), ("bf16", capi.OPF_BF16))
    print("| weights | seed | " + " | ".join(f"{m}: wave max-abs / worst stage rel-L2 (stage)" for m, _ in modes) + " | wave peak |")
    print("|---|---|" + "---|" * (len(modes) + 1))
    worst = {m: [0.0, 0.0] for m, _ in modes}
    for family in ("synth", "ctor"):
        for seed in range(n):
            sd = synth.synthetic_state_dict(shapes, seed) if family == "synth" else ctor_state_dict(cfg, seed)
            sd = {k: v for k, v in sd.items()}
            unit, mel, noise = synth.synthetic_inputs(1, 250, 1, 250, 10 + seed)
            ref = {}
            qvc_oracle.infer(sd, unit, mel, noise, dtype=torch.float64, taps=ref)
            cells = []
            for name, opf in modes:
                taps = {}
                wave = Emu(sd, opf).infer(unit, mel, noise, taps)
                errs = {t: synth.rel_l2(taps[t], ref[t]) for t in qvc_oracle.TAP_NAMES}
                ws = max(errs, key=errs.get)
                e = synth.max_abs(wave, ref["wave"])
                worst[name][0] = max(worst[name][0], e)
                worst[name][1] = max(worst[name][1], errs[ws])
                cells.append(f"{e:.2e} / {errs[ws]:.2e} ({ws})")
            print(f"| {family} | {seed} | " + " | ".join(cells) + f" | {float(ref['wave'].abs().max()):.3f} |", flush=True)
    print()
    for name, _ in modes:
        print(f"worst over all rows, {name}: waveform max-abs {worst[name][0]:.2e}, per-stage rel-L2 {worst[name][1]:.2e}")


if __name__ == "__main__":
    main()
