"""PipelinedConverter on a B200: same waveforms as direct infer calls, in order, with buffer reuse."""
import json
import os

import pytest
import torch

import synth
from conftest import ROOT
from quickvc_official_b200 import SynthesizerTrn
from quickvc_official_b200.pipeline import GraphedInfer, PipelinedConverter

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_pipelined_converter_matches_direct_calls():
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "quickvc_model_config.json")))
    shapes = {k: tuple(v) for k, v in json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json"))).items()}
    sd = synth.synthetic_state_dict(shapes, 0)
    net = SynthesizerTrn(641, 32, **cfg).eval()
    net.load_state_dict(sd)
    net = net.to(DEV)
    B, T, TM, n = 3, 40, 150, 5
    batches, want = [], []
    for i in range(n):
        unit, mel, _ = synth.synthetic_inputs(B, T, 1, TM, 20 + i)
        batches.append((unit.pin_memory(), mel.pin_memory()))
    # infer draws its own noise: fix the generator so both passes see the same draws
    torch.manual_seed(123)
    torch.cuda.manual_seed(123)
    for unit, mel in batches:
        want.append(net.infer(unit.to(DEV), mel.to(DEV)).cpu())
    torch.manual_seed(123)
    torch.cuda.manual_seed(123)
    conv = PipelinedConverter(net, B, T, TM)
    got = list(conv.convert_many(batches))
    assert len(got) == n
    for g, w in zip(got, want):
        assert g.shape == (B, 1, 320 * T) and torch.equal(g, w)


def test_graphed_infer_replays_the_same_arithmetic():
    """A call captured into a CUDA graph (speaker-encoder fork / join included) gives the eager call's bits on new inputs."""
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "quickvc_model_config.json")))
    shapes = {k: tuple(v) for k, v in json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_shapes.json"))).items()}
    net = SynthesizerTrn(641, 32, **cfg).eval()
    net.load_state_dict(synth.synthetic_state_dict(shapes, 0))
    net = net.to(DEV)
    B, T, TM = 2, 50, 200
    graphed = GraphedInfer(net, B, T, mel_frames=TM)
    cached = GraphedInfer(net, B, T)
    for seed in (1, 2, 3):
        unit, mel, noise = (t.to(DEV) for t in synth.synthetic_inputs(B, T, 1, TM, seed))
        want = net.infer(unit, mel, noise=noise)
        got = graphed(unit, mel, noise)
        assert torch.equal(got, want)
        g = net.embed_speaker(mel)
        assert torch.equal(cached(unit, g, noise), net.infer_with_embedding(unit, g, noise=noise))
    # without a supplied noise the draw happens inside the static buffer: finite output of the right shape
    out = graphed(unit, mel)
    assert out.shape == (B, 1, 320 * T) and bool(torch.isfinite(out).all())


def test_host_gather_equals_infer(sd, model_cfg):
    """shard.HostGather on one GPU: the shared, page-locked host buffer receives exactly net.infer's waveforms, chunk by
    chunk on the side stream (the N-rank form is covered on gloo in tests/test_shard.py)."""
    from quickvc_official_b200.shard import HostGather
    net = SynthesizerTrn(641, 32, **model_cfg, precision="tf32").eval()
    net.load_state_dict(sd)
    net = net.to(DEV)
    n, t = 5, 40
    unit, mel, noise = synth.synthetic_inputs(n, t, 1, 150, 3)
    u, m = unit.to(DEV), mel.to(DEV)
    torch.manual_seed(0)
    want = net.infer(u, m, noise=noise.to(DEV))
    hg = HostGather(n, 320 * t, name="qvc_test_gather_gpu")
    try:
        nz = noise.to(DEV)
        pos = [0]

        def infer(uu, mm):
            k = uu.shape[0]
            out = net.infer(uu, mm, noise=nz[pos[0]:pos[0] + k])
            pos[0] += k
            return out

        got = hg.convert(infer, u, m, chunk=2)
        hg.finish(DEV)
        assert got.shape == (n, 1, 320 * t) and not got.is_cuda
        assert torch.equal(got, want.cpu())
    finally:
        hg.close()
