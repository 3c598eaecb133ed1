"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in-process.

Run in the authoring container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Recipe: scipy.signal.kaiser shim (pqmf.py:13 imports a name SciPy >= 1.13
removed; PQMF itself is never instantiated under ms_istft_vits), build SynthesizerTrn(641, 32,
**model-config), load the synthetic state_dict of tests/synth.py, patch torch.randn_like to return the
injected noise (models.py:94), hook the stage taps of SURVEY.md section 8a, call `infer`.
"""
import json
import os
import sys

import numpy as np
import scipy.signal
import scipy.signal.windows
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("QVC_REFERENCE", "/root/reference")

import synth  # noqa: E402

CASES = {
    # name: (B, T, Bm, Tm, taps kept)
    "small":    (2, 24, 1, 200, "all"),
    "shortmel": (2, 16, 2, 100, ("g", "z_p", "flow_0", "wave")),
    "cfg1":     (1, 250, 1, 250, ("g", "wave")),
    "chunk":    (1, 25, 1, 250, ("wave",)),
}
SEED = 0


def load_reference():
    scipy.signal.kaiser = scipy.signal.windows.kaiser
    sys.path.insert(0, REF)
    import models  # the reference's models.py
    return models


def run_reference(net, unit, mel, noise):
    taps = {}
    hooks = []

    def keep(name, fn=lambda o: o):
        def hook(_m, _i, o):
            taps[name] = fn(o).detach().clone()
        return hook

    def enc_p_hook(_m, _i, o):
        taps["z_p"], taps["m_p"], taps["logs_p"] = (t.detach().clone() for t in o)
    hooks.append(net.enc_p.register_forward_hook(enc_p_hook))
    for idx in (6, 4, 2, 0):
        hooks.append(net.flow.flows[idx].register_forward_hook(keep(f"flow_{idx}")))
    hooks.append(net.dec.conv_pre.register_forward_hook(keep("_conv_pre")))
    hooks.append(net.dec.cond.register_forward_hook(keep("_cond")))
    for i in range(2):
        hooks.append(net.dec.ups[i].register_forward_hook(keep(f"ups_{i}")))
    for r in range(6):
        hooks.append(net.dec.resblocks[r].register_forward_hook(keep(f"_rb{r}")))
    hooks.append(net.dec.subband_conv_post.register_forward_hook(keep("conv_post")))
    hooks.append(net.dec.stft.register_forward_hook(keep("_istft")))

    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: noise.to(t.dtype)
    try:
        with torch.no_grad():
            wave = net.infer(unit, mel)
            g = net.enc_spk.embed_utterance(mel.transpose(1, 2)).unsqueeze(-1)
    finally:
        torch.randn_like = orig
        for h in hooks:
            h.remove()
    taps["g"] = g
    taps["conv_pre"] = taps.pop("_conv_pre") + taps.pop("_cond")
    for i in range(2):      # models.py:378-384: xs = rb0; xs += rb1; xs += rb2; x = xs / 3
        taps[f"mrf_{i}"] = ((taps.pop(f"_rb{3 * i}") + taps.pop(f"_rb{3 * i + 1}")) + taps.pop(f"_rb{3 * i + 2}")) / 3
    b = unit.shape[0]
    taps["y_mb"] = taps.pop("_istft").reshape(b, 4, -1)
    taps["wave"] = wave
    return taps


def build_reference_net(seed=SEED):
    models = load_reference()
    cfg = json.load(open(os.path.join(HERE, "quickvc_model_config.json")))
    net = models.SynthesizerTrn(641, 32, **cfg).eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = synth.synthetic_state_dict(shapes, seed)
    net.load_state_dict(sd)
    return net, shapes, sd


def main():
    net, shapes, sd = build_reference_net()
    with open(os.path.join(HERE, "state_dict_shapes.json"), "w") as f:
        json.dump({k: list(v) for k, v in shapes.items()}, f, indent=0)
    for name, (b, t, bm, tm, keep) in CASES.items():
        unit, mel, noise = synth.synthetic_inputs(b, t, bm, tm, SEED)
        taps = run_reference(net, unit, mel, noise)
        if keep != "all":
            taps = {k: v for k, v in taps.items() if k in keep}
        out = os.path.join(HERE, f"infer_{name}.npz")
        np.savez_compressed(out, **{k: v.numpy().astype(np.float32) for k, v in taps.items()})
        w = taps["wave"]
        print(f"{name}: B={b} T={t} Bm={bm} Tm={tm} -> wave {tuple(w.shape)} rms {w.pow(2).mean().sqrt():.4f} "
              f"peak {w.abs().max():.4f}  [{os.path.getsize(out) / 1e6:.2f} MB]")


if __name__ == "__main__":
    main()
