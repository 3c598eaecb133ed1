"""Deterministic synthetic weights and inputs shared by the golden generator and the tests.

Values come from numpy's Philox generator keyed by (seed, crc32(name)), so they are bit-identical on
every machine and independent of torch's RNG, of module construction order and of the reference.
Scales mimic torch's default initialisation (the reference's actual init);
`weight_g` is perturbed away from ||v|| so the weight-norm fold is exercised, and the flow `post`
layers are non-zero (the reference zero-inits them, modules.py:196-197, which would make the flow an
identity and parity vacuous).
"""
from __future__ import annotations

import zlib
from typing import Dict, Mapping, Tuple

import numpy as np
import torch


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed, zlib.crc32(name.encode())]))


def synthetic_state_dict(shapes: Mapping[str, Tuple[int, ...]], seed: int = 0,
                         buffers: Mapping[str, torch.Tensor] | None = None) -> Dict[str, torch.Tensor]:
    """shapes: the 467 (key -> shape) pairs of the reference layout.  Buffers keep their fixed values."""
    out: Dict[str, torch.Tensor] = {}
    for key, shape in shapes.items():
        if key in ("dec.updown_filter", "dec.stft.window"):
            continue
        if key.endswith("weight_g"):
            continue
        r = _rng(seed, key)
        shape = tuple(shape)
        if "lstm" in key:
            bound = 1.0 / 16.0
        elif key.endswith(("weight_v", "weight")):
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            bound = 1.0 / np.sqrt(fan_in)
        else:  # biases
            bound = 0.05
        out[key] = torch.from_numpy(r.uniform(-bound, bound, size=shape).astype(np.float32))
    for key, shape in shapes.items():
        if key.endswith("weight_g"):
            v = out[key[:-1] + "v"]
            nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(tuple(shape))
            scale = torch.from_numpy(_rng(seed, key).uniform(0.7, 1.3, size=tuple(shape)).astype(np.float32))
            out[key] = nrm * scale
    updown = torch.zeros(4, 4, 4)
    for k in range(4):
        updown[k, k, 0] = 1.0
    out["dec.updown_filter"] = updown if buffers is None else buffers["dec.updown_filter"].clone()
    out["dec.stft.window"] = torch.hann_window(16) if buffers is None else buffers["dec.stft.window"].clone()
    return {k: out[k] for k in shapes}       # reference key order


def synthetic_inputs(batch: int, frames: int, mel_batch: int, mel_frames: int, seed: int = 0):
    """unit (B,256,T) ~ N(0,1); mel (Bm,80,Tm) ~ log-mel-like N(-5, 2); noise (B,192,T) ~ N(0,1)."""
    unit = torch.from_numpy(_rng(seed, "unit").standard_normal((batch, 256, frames)).astype(np.float32))
    mel = torch.from_numpy((_rng(seed, "mel").standard_normal((mel_batch, 80, mel_frames)) * 2.0 - 5.0).astype(np.float32))
    noise = torch.from_numpy(_rng(seed, "noise").standard_normal((batch, 192, frames)).astype(np.float32))
    return unit, mel, noise


def reference_shapes() -> Dict[str, Tuple[int, ...]]:
    """The reference layout, taken from our drop-in module (its equality with the reference's own
    state_dict is asserted in tests/test_boundary.py whenever /root/reference is present)."""
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from quickvc_official_b200 import SynthesizerTrn
    cfg = json.load(open(os.path.join(root, "tests", "golden", "quickvc_model_config.json")))
    net = SynthesizerTrn(641, 32, **cfg)
    return {k: tuple(v.shape) for k, v in net.state_dict().items()}


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double().cpu() - b.double().cpu()).abs().max())
