"""Ragged batches (`lengths`): every utterance of a padded batch gets exactly the waveform it gets alone.

The reference has no batching of unequal lengths -- convert.py:59-86 converts one utterance per call -- so the
contract is defined by the single-utterance call: wave[b, :, :320 len_b] == infer(unit[b, :, :len_b]) and zeros after,
whatever the padding frames of `unit` / `noise` hold.  Checked against our own single-utterance calls (bit for bit:
the per-element arithmetic does not depend on the tiling) and against the fp64 CPU oracle.
"""
import numpy as np
import pytest
import torch

import synth
from oracle import qvc_oracle
from quickvc_official_b200 import SynthesizerTrn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _net(sd, model_cfg, precision):
    net = SynthesizerTrn(641, 32, precision=precision, **model_cfg).eval()
    net.load_state_dict(sd)
    return net.to(DEV)


def _ragged_inputs(B, T, lens, seed):
    unit, mel, noise = synth.synthetic_inputs(B, T, 1, 200, seed)
    junk = torch.from_numpy(np.random.default_rng(seed).standard_normal((B, 256, T)).astype(np.float32)) * 50.0
    for b, n in enumerate(lens):            # padding frames hold large junk: it must not reach the live samples
        unit[b, :, n:] = junk[b, :, n:]
        noise[b, :, n:] = 7.0
    return unit.to(DEV), mel.to(DEV), noise.to(DEV)


@pytest.mark.parametrize("precision,tol", [("tf32", 0.0), ("bf16", 0.0), ("fp16", 0.0), ("fp32", 0.0)])
def test_small_ragged_batch_equals_single_calls(sd, model_cfg, precision, tol):
    net = _net(sd, model_cfg, precision)
    T = 96
    lens = [96, 37, 64, 1, 33, 65, 5]
    B = len(lens)
    unit, mel, noise = _ragged_inputs(B, T, lens, 3)
    wave = net.infer(unit, mel, noise=noise, lengths=torch.tensor(lens))
    assert wave.shape == (B, 1, 320 * T)
    for b, n in enumerate(lens):
        alone = net.infer(unit[b:b + 1, :, :n].contiguous(), mel, noise=noise[b:b + 1, :, :n].contiguous())
        got = wave[b, 0, :320 * n]
        diff = float((got - alone[0, 0]).abs().max())
        assert diff <= tol, f"{precision}: utterance {b} (len {n}) differs from its single call by {diff}"
        assert float(wave[b, 0, 320 * n:].abs().max()) == 0.0 if n < T else True


def test_ragged_batch_on_pair_and_fused_kernels(sd, model_cfg):
    """Enough tiles for the CTA-pair and fused-WN kernels (the B = 64 x 10 s code path), lengths on and off tile edges."""
    net = _net(sd, model_cfg, "tf32")
    T, B = 256, 40
    rng = np.random.default_rng(11)
    lens = [int(x) for x in rng.integers(40, T + 1, B)]
    lens[0], lens[1], lens[2], lens[3] = T, 128, 129, 255
    unit, mel, noise = _ragged_inputs(B, T, lens, 4)
    lens_dev = torch.tensor(lens, dtype=torch.int64, device=DEV)       # a device tensor is taken as is
    wave = net.infer(unit, mel, noise=noise, lengths=lens_dev)
    for b in (0, 1, 2, 3, 17, 39):
        n = lens[b]
        alone = net.infer(unit[b:b + 1, :, :n].contiguous(), mel, noise=noise[b:b + 1, :, :n].contiguous())
        diff = float((wave[b, 0, :320 * n] - alone[0, 0]).abs().max())
        assert diff <= 2e-6, f"utterance {b} (len {n}) differs from its single call by {diff}"
        if n < T:
            assert float(wave[b, 0, 320 * n:].abs().max()) == 0.0
    # and against the fp64 oracle on one short utterance (fp32-mode tolerance of BASELINE.json's north_star)
    b, n = 17, lens[17]
    ref = qvc_oracle.infer(sd, unit[b:b + 1, :, :n].cpu(), mel.cpu(), noise[b:b + 1, :, :n].cpu(), dtype=torch.float64)
    assert float((wave[b, 0, :320 * n].cpu().double() - ref[0, 0]).abs().max()) <= 1e-4


def test_ragged_with_per_utterance_embeddings_and_decode(sd, model_cfg):
    net = _net(sd, model_cfg, "tf32")
    T, lens = 80, [80, 23, 51]
    B = len(lens)
    unit, mel, noise = _ragged_inputs(B, T, lens, 5)
    mels = [synth.synthetic_inputs(1, 1, 1, 150 + 30 * b, 40 + b)[1].to(DEV) for b in range(B)]
    g = torch.cat([net.embed_speaker(m) for m in mels], dim=0)
    wave = net.infer_with_embedding(unit, g, noise=noise, lengths=torch.tensor(lens))
    for b, n in enumerate(lens):
        alone = net.infer(unit[b:b + 1, :, :n].contiguous(), mels[b], noise=noise[b:b + 1, :, :n].contiguous())
        assert float((wave[b, 0, :320 * n] - alone[0, 0]).abs().max()) <= 2e-6
    # decoder-only entry point
    z = torch.randn(B, 192, T, generator=torch.Generator().manual_seed(1)).to(DEV)
    for b, n in enumerate(lens):
        z[b, :, n:] = 99.0
    wd = net.decode(z, g.unsqueeze(-1), lengths=torch.tensor(lens))
    for b, n in enumerate(lens):
        alone = net.decode(z[b:b + 1, :, :n].contiguous(), g[b:b + 1].unsqueeze(-1))
        assert float((wd[b, 0, :320 * n] - alone[0, 0]).abs().max()) <= 2e-6
        if n < T:
            assert float(wd[b, 0, 320 * n:].abs().max()) == 0.0


def test_lengths_argument_checks(sd, model_cfg):
    net = _net(sd, model_cfg, "tf32")
    unit, mel, noise = _ragged_inputs(2, 16, [16, 16], 6)
    with pytest.raises(ValueError):
        net.infer(unit, mel, noise=noise, lengths=torch.tensor([16]))
    with pytest.raises(ValueError):
        net.infer(unit, mel, noise=noise, lengths=torch.tensor([16, 17]))
    with pytest.raises(ValueError):
        net.infer(unit, mel, noise=noise, lengths=torch.tensor([0, 5]))
    with pytest.raises(ValueError):
        net.infer(unit, mel, noise=noise, lengths=torch.tensor([4.0, 5.0]))
    full = net.infer(unit, mel, noise=noise)
    same = net.infer(unit, mel, noise=noise, lengths=torch.tensor([16, 16]))
    assert torch.equal(full, same)


@pytest.mark.parametrize("precision", ["tf32", "fp16"])
def test_dead_tiles_are_skipped_without_reading_stale_memory(sd, model_cfg, precision):
    """Tiles wholly past an utterance's end are skipped by every warp role (DEAD_MARGIN, csrc/common.cuh), so their rows of
    the workspace keep whatever an earlier call left there.  Poison the workspace with NaN bit patterns first: nothing of
    it may reach a live sample, and the step must get cheaper than the dense one."""
    net = _net(sd, model_cfg, precision)
    T, B = 640, 24
    lens = [T, 40, 300, 33, 257, 1, 512, 129] * 3
    unit, mel, noise = _ragged_inputs(B, T, lens, 9)
    lens_t = torch.tensor(lens)
    net.infer(unit, mel, noise=noise, lengths=lens_t)                       # sizes the workspace
    for ws in net._engine._ws.values():
        ws.fill_(0xFF)                                                      # every fp32 / 16-bit word a NaN
    wave = net.infer(unit, mel, noise=noise, lengths=lens_t)
    assert bool(torch.isfinite(wave).all())
    for b in (0, 1, 2, 3, 4, 5, 6, 7, 23):
        n = lens[b]
        alone = net.infer(unit[b:b + 1, :, :n].contiguous(), mel, noise=noise[b:b + 1, :, :n].contiguous())
        diff = float((wave[b, 0, :320 * n] - alone[0, 0]).abs().max())
        assert diff <= 2e-6, f"utterance {b} (len {n}) differs from its single call by {diff}"
        if n < T:
            assert float(wave[b, 0, 320 * n:].abs().max()) == 0.0

    def ms(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    dense = ms(lambda: net.infer(unit, mel, noise=noise))
    ragged = ms(lambda: net.infer(unit, mel, noise=noise, lengths=lens_t))
    live = sum(lens) / (B * T)
    print(f"{precision}: dense {dense:.3f} ms, ragged {ragged:.3f} ms, live fraction {live:.2f}")
    assert ragged < dense * 0.9          # dead tiles cost nothing; the live ones of the longest utterances set the critical path
