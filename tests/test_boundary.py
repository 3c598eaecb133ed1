"""Drop-in boundary: constructor contract, state_dict layout, C-ABI exports (no GPU needed)."""
import ctypes
import os
import re
import sys

import pytest
import torch

from conftest import REFERENCE, ROOT
from quickvc_official_b200 import SynthesizerTrn, capi


def test_state_dict_layout_matches_reference(shapes, model_cfg):
    net = SynthesizerTrn(641, 32, **model_cfg)
    got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert len(got) == 467
    assert list(got) == list(shapes)           # same keys in the same order
    assert got == shapes
    n_inf = sum(v.numel() for k, v in net.state_dict().items() if not k.startswith("enc_q."))
    assert sum(p.numel() for p in net.parameters()) == 40032285 - 16 - 64   # buffers are not parameters
    assert n_inf > 31e6


def test_load_state_dict_roundtrip(sd, model_cfg):
    net = SynthesizerTrn(641, 32, **model_cfg)
    net.load_state_dict(sd)                    # strict
    for k, v in net.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_constructor_contract(model_cfg):
    bad = dict(model_cfg, ms_istft_vits=False)
    with pytest.raises(RuntimeError):          # models.py:589
        SynthesizerTrn(641, 32, **bad)
    with pytest.raises(AssertionError):        # models.py:574-575
        SynthesizerTrn(641, 32, **model_cfg, resblock="2")
    SynthesizerTrn(641, 32, **model_cfg, resblock="1", n_heads=2, p_dropout=0.1)   # legacy keys are swallowed
    with pytest.raises(NotImplementedError):
        SynthesizerTrn(641, 32, **dict(model_cfg, hidden_channels=128))
    with pytest.raises(ValueError):
        SynthesizerTrn(641, 32, **model_cfg, precision="fp8")


def test_infer_has_no_cpu_fallback(sd, model_cfg):
    net = SynthesizerTrn(641, 32, **model_cfg).eval()
    with pytest.raises(capi.QvcError):
        net.infer(torch.zeros(1, 256, 8), torch.zeros(1, 80, 130))


def test_forward_is_out_of_scope(model_cfg):
    net = SynthesizerTrn(641, 32, **model_cfg)
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 256, 8), torch.zeros(1, 641, 8), torch.zeros(1, 80, 8))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "qvc_b200.h")).read()
    declared = set(re.findall(r"\b(qvc_[a-z0-9_]+)\s*\(", header))
    declared -= {"qvc_layer_index_"}
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    lib = capi.load()                          # dlopen + bind all; raises on a missing symbol
    assert lib.qvc_abi_version() == capi.QVC_ABI_VERSION
    raw = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name


def test_struct_sizes_match_header():
    # sizes the C compiler produces for the same declarations (LP64)
    assert ctypes.sizeof(capi.Tensor) == 24
    assert ctypes.sizeof(capi.EpiSegment) == 24 + 5 * 24 + 8
    assert ctypes.sizeof(capi.Layer) == 40
    assert ctypes.sizeof(capi.ConvArgs) == 24 + 16 + 24 + 16 + 8 + 2 * 152 + 3 * 24 + 8 + 16 + 32
    assert ctypes.sizeof(capi.Model) == 16 + (2 * 114 + 5) * 40 + 24 + 11 * 8 + 4 * 8


def test_integration_md_binding_stub_is_valid_python_against_the_real_binding():
    """The ctypes stub INTEGRATION.md shows a maintainer compiles, and every `_lib.` / `capi.` name it uses exists."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, re.S)
    stub = next(b for b in blocks if "class B200Infer" in b)
    compile(stub, "INTEGRATION.md", "exec")
    for sym in set(re.findall(r"_lib\.(qvc_[a-z0-9_]+)", stub)):
        assert sym in capi.SYMBOLS, sym
    for attr in set(re.findall(r"capi\.([A-Za-z_][A-Za-z0-9_]*)", stub)):
        assert hasattr(capi, attr), attr
    # the argument counts of the two calls it makes match the bound prototypes
    for sym in ("qvc_prepare_weights", "qvc_infer"):
        call = re.search(sym + r"\((.*?)\)(?:,\s*\"|\n\s+capi\.check)", stub, re.S).group(1)
        depth, n = 0, 1
        for ch in call:
            depth += ch in "([" ; depth -= ch in ")]"
            n += ch == "," and depth == 0
        assert n == len(capi.SYMBOLS[sym][1]), (sym, n, len(capi.SYMBOLS[sym][1]))


def test_plain_c_consumer_binds_the_abi(tmp_path):
    """include/qvc_b200.h is a C header (C99, -pedantic -Werror) and the library is usable from plain C with no Python, torch
    or CUDA headers: tests/c/consumer.c dlopens it, checks the ABI version and size queries, and gets the fold's QVC_ERR_ARG
    + error string for a misshapen state_dict entry.  The struct sizes the C compiler derives from the header are the ones
    the ctypes mirror (capi.py) uses."""
    import subprocess
    exe = str(tmp_path / "consumer")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "consumer.c"), "-o", exe, "-ldl"], check=True)
    r = subprocess.run([exe, capi.LIB_PATH], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert f"ok abi={capi.QVC_ABI_VERSION} " in r.stdout and "enc_p.pre.weight" in r.stdout
    sizes = dict(kv.split("=") for kv in next(l for l in r.stdout.splitlines() if l.startswith("sizeof ")).split()[1:])
    mirror = {"qvc_tensor": capi.Tensor, "qvc_epi_segment": capi.EpiSegment, "qvc_conv_args": capi.ConvArgs,
              "qvc_spk_weights": capi.SpkWeights, "qvc_mel_weights": capi.MelWeights, "qvc_tail_weights": capi.TailWeights,
              "qvc_layer": capi.Layer, "qvc_model": capi.Model, "qvc_state_entry": capi.StateEntry, "qvc_taps": capi.Taps}
    for name, cls in mirror.items():
        assert int(sizes[name]) == ctypes.sizeof(cls), (name, sizes[name], ctypes.sizeof(cls))


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not mounted")
def test_seeded_init_equals_reference(model_cfg):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    models = make_golden.load_reference()
    torch.manual_seed(0)
    ref = models.SynthesizerTrn(641, 32, **model_cfg).state_dict()
    torch.manual_seed(0)
    ours = SynthesizerTrn(641, 32, **model_cfg).state_dict()
    assert list(ref) == list(ours)
    for k in ref:
        assert torch.equal(ref[k], ours[k]), k
