"""The CPU oracle (oracle/qvc_oracle.py) against the fixtures the reference produced, and -- when the
reference tree is mounted -- against the reference itself."""
import os
import sys

import pytest
import torch

import synth
from conftest import GOLDEN_CASES, REFERENCE, load_golden
from oracle import qvc_oracle


@pytest.mark.parametrize("case", list(GOLDEN_CASES))
def test_oracle_matches_reference_golden(case, sd):
    b, t, bm, tm = GOLDEN_CASES[case]
    unit, mel, noise = synth.synthetic_inputs(b, t, bm, tm, 0)
    gold = load_golden(case)
    taps = {}
    wave = qvc_oracle.infer(sd, unit, mel, noise, taps=taps)
    assert wave.shape == (b, 1, 320 * t)
    for name, ref in gold.items():
        got = taps[name]
        assert got.shape == ref.shape, name
        # fp32 vs fp32 with a different summation order (closed-form iSTFT, explicit LSTM loop)
        assert synth.rel_l2(got, ref) < 2e-5, (name, synth.rel_l2(got, ref))
    assert synth.max_abs(wave, gold["wave"]) < 2e-6


def test_oracle_fp64_close_to_fp32(sd):
    b, t, bm, tm = GOLDEN_CASES["small"]
    unit, mel, noise = synth.synthetic_inputs(b, t, bm, tm, 0)
    w32 = qvc_oracle.infer(sd, unit, mel, noise)
    w64 = qvc_oracle.infer(sd, unit, mel, noise, dtype=torch.float64)
    assert synth.max_abs(w32, w64) < 2e-6


def test_oracle_batched_long_mel_rejected(sd):
    unit, mel, noise = synth.synthetic_inputs(2, 8, 2, 200, 0)
    with pytest.raises(ValueError):
        qvc_oracle.infer(sd, unit, mel, noise)


def test_window_starts():
    # models.py:520-535: range(0, Tm-128, 64) plus the last 128 frames
    assert qvc_oracle.window_starts(129) == [0, 1]
    assert qvc_oracle.window_starts(250) == [0, 64, 122]
    assert len(qvc_oracle.window_starts(500)) == 7


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not mounted")
def test_oracle_matches_live_reference(sd):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    net, _, sd_ref = make_golden.build_reference_net()
    unit, mel, noise = synth.synthetic_inputs(1, 40, 1, 130, 3)
    ref = make_golden.run_reference(net, unit, mel, noise)
    taps = {}
    qvc_oracle.infer(sd_ref, unit, mel, noise, taps=taps)
    for name in qvc_oracle.TAP_NAMES:
        assert synth.rel_l2(taps[name], ref[name]) < 2e-5, name
