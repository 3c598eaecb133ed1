"""Size-independent properties at BASELINE.json's full sizes (configs[1]: 64 x 10 s), on a B200.

The oracle needs ~30 s of CPU per such batch, so the full-size checks use properties instead:
  * utterances are independent: any utterance of the big batch equals, bit for bit,
    the same utterance converted alone -- tiles never leak across utterance boundaries, whatever the tiling;
  * chunked decoding (chunk_utts) does not change a single bit;
  * one utterance of the big batch is checked against the CPU oracle with the fp32-mode tolerance.
"""
import json
import os

import pytest
import torch

import synth
from conftest import ROOT
from oracle import qvc_oracle
from quickvc_official_b200 import SynthesizerTrn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def big():
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "quickvc_model_config.json")))
    shapes = {k: tuple(v) for k, v in json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json"))).items()}
    sd = synth.synthetic_state_dict(shapes, 0)
    unit, mel, noise = synth.synthetic_inputs(64, 500, 1, 500, 11)
    return cfg, sd, unit, mel, noise


@pytest.mark.parametrize("precision", ["tf32", "bf16", "fp16"])
def test_utterances_are_independent_at_full_size(big, precision):
    cfg, sd, unit, mel, noise = big
    net = SynthesizerTrn(641, 32, **cfg, precision=precision).eval()
    net.load_state_dict(sd)
    net = net.to(DEV)
    u, m, n = unit.to(DEV), mel.to(DEV), noise.to(DEV)
    wave = net.infer(u, m, noise=n)
    assert wave.shape == (64, 1, 160000) and bool(torch.isfinite(wave).all())
    for b in (0, 37, 63):
        alone = net.infer(u[b:b + 1], m, noise=n[b:b + 1])
        assert torch.equal(alone[0], wave[b]), f"utterance {b} differs between batch 64 and batch 1"
    # a different sub-batching of the decoder is the same arithmetic
    net_c = SynthesizerTrn(641, 32, **cfg, precision=precision, chunk_utts=24).eval()
    net_c.load_state_dict(sd)
    net_c = net_c.to(DEV)
    assert torch.equal(net_c.infer(u, m, noise=n), wave)


def test_one_full_length_utterance_against_oracle(big):
    cfg, sd, unit, mel, noise = big
    net = SynthesizerTrn(641, 32, **cfg, precision="tf32").eval()
    net.load_state_dict(sd)
    net = net.to(DEV)
    wave = net.infer(unit.to(DEV), mel.to(DEV), noise=noise.to(DEV))
    b = 5
    ref = qvc_oracle.infer(sd, unit[b:b + 1], mel, noise[b:b + 1])
    err = float((wave[b:b + 1].cpu() - ref).abs().max())
    assert err < 1e-4, err            # north-star fp32-mode bound: max-abs waveform error 1e-4


def test_one_minute_utterance_and_a_ragged_companion(big):
    """A single 60 s utterance (T = 3000: 60000-row series in the decoder, 235 tiles per layer) against the oracle, and
    the same utterance in a ragged batch next to a 7 s one."""
    cfg, sd, _, mel, _ = big
    net = SynthesizerTrn(641, 32, **cfg, precision="tf32").eval()
    net.load_state_dict(sd)
    net = net.to(DEV)
    T = 3000
    unit, _, noise = synth.synthetic_inputs(2, T, 1, 8, 21)
    u, m, n = unit.to(DEV), mel.to(DEV), noise.to(DEV)
    long_alone = net.infer(u[:1], m, noise=n[:1])
    assert long_alone.shape == (1, 1, 320 * T)
    ref = qvc_oracle.infer(sd, unit[:1], mel, noise[:1])
    assert float((long_alone.cpu() - ref).abs().max()) < 1e-4
    lens = torch.tensor([T, 350])
    both = net.infer(u, m, noise=n, lengths=lens)
    assert float((both[0] - long_alone[0]).abs().max()) <= 2e-6
    short_alone = net.infer(u[1:2, :, :350].contiguous(), m, noise=n[1:2, :, :350].contiguous())
    assert float((both[1, 0, :320 * 350] - short_alone[0, 0]).abs().max()) <= 2e-6
    assert float(both[1, 0, 320 * 350:].abs().max()) == 0.0
