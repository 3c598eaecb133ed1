"""qvc_fold_host / qvc_prepare_weights (csrc/fold.cu, the fold a non-Python consumer of the C ABI uses) against the Python
statement of the same fold (quickvc-official_b200/fold.py), tensor by tensor, on the CPU."""
import ctypes as C

import numpy as np
import pytest
import torch

from quickvc_official_b200 import capi, fold

_DT = {capi.OPF_F32: torch.float32, capi.OPF_TF32: torch.float32, capi.OPF_BF16: torch.bfloat16, capi.OPF_F16: torch.float16}


def _native(sd, opf, backend=capi.BACKEND_FMA):
    lib = capi.load()
    entries, keep = capi.state_entries(sd)
    nbytes = int(lib.qvc_prepared_bytes(opf))
    assert nbytes > 0
    block = torch.zeros(nbytes, dtype=torch.uint8)
    tail_host = torch.zeros(16 + 272, dtype=torch.float32)
    model = capi.Model()
    capi.check(lib.qvc_fold_host(entries, len(entries), opf, backend, block.data_ptr(), nbytes, tail_host.data_ptr(),
                                 C.byref(model)), "qvc_fold_host")
    return model, block, tail_host, keep


def _view(block, ptr, numel, dtype):
    off = ptr - block.data_ptr()
    assert 0 <= off and off + numel * dtype.itemsize <= block.numel()
    return block[off: off + numel * dtype.itemsize].view(dtype)


@pytest.mark.parametrize("opf", [capi.OPF_F32, capi.OPF_TF32, capi.OPF_BF16, capi.OPF_F16], ids=["f32", "tf32", "bf16", "f16"])
def test_native_fold_equals_python_fold(sd, opf):
    f = fold.fold_state_dict({k: v for k, v in sd.items() if not k.startswith("enc_q.")}, opf)
    model, block, tail_host, _keep = _native(sd, opf)
    assert model.abi_version == capi.QVC_ABI_VERSION and model.opformat == opf and model.cond_rows == fold.COND_ROWS
    worst = 0
    for i, L in enumerate(f.layers):
        M = model.layers[i]
        assert (M.cin, M.cout, M.k, M.dil, M.pad_left) == (L["cin"], L["cout"], L["k"], L["dil"], L["pad_left"]), L["name"]
        got = _view(block, M.w, L["w"].numel(), _DT[opf])
        want = L["w"].reshape(-1)
        # identical arithmetic up to the order of the fp64 sum under the weight-norm square root: at most one unit of the
        # operand format on a vanishing fraction of the elements
        same = got.view(torch.int16 if _DT[opf].itemsize == 2 else torch.int32) == want.view(torch.int16 if _DT[opf].itemsize == 2 else torch.int32)
        frac = float((~same).float().mean())
        assert frac < 1e-4, (L["name"], frac)
        assert torch.allclose(got.float(), want.float(), rtol=2 ** -7 if opf == capi.OPF_BF16 else 2 ** -9, atol=1e-30), L["name"]
        worst = max(worst, frac)
        if L["bias"] is None:
            assert not M.bias
        else:
            assert torch.equal(_view(block, M.bias, L["bias"].numel(), torch.float32), L["bias"]), L["name"]
    for i in range(capi.QVC_NUM_LAYERS):
        P = model.paired[i]
        if i not in f.paired:
            assert not P.w
            continue
        L = f.paired[i]
        assert (P.cin, P.cout, P.k, P.dil, P.pad_left) == (L["cin"], L["cout"], L["k"], L["dil"], L["pad_left"])
        got = _view(block, P.w, L["w"].numel(), _DT[opf])
        assert float((got.float() != L["w"].reshape(-1).float()).float().mean()) < 1e-4
        assert torch.equal(_view(block, P.bias, L["bias"].numel(), torch.float32), L["bias"])
    assert len(f.wn_skip) == 5
    for i, L in enumerate(f.wn_skip):
        S = model.wn_skip[i]
        assert (S.cin, S.cout, S.k, S.dil, S.pad_left) == (L["cin"], L["cout"], 1, 1, 0)
        got = _view(block, S.w, L["w"].numel(), _DT[opf])
        assert float((got.float() != L["w"].reshape(-1).float()).float().mean()) < 1e-4, i
        assert torch.allclose(_view(block, S.bias, 192, torch.float32), L["bias"], rtol=1e-6, atol=1e-7), i
    t = f.tensors
    assert torch.equal(_view(block, model.cond_w, t["cond_w"].numel(), torch.float32), t["cond_w"].reshape(-1))
    assert torch.equal(_view(block, model.cond_b, t["cond_b"].numel(), torch.float32), t["cond_b"])
    for l in range(3):
        for name, ptr in (("w_ih", model.spk.w_ih[l]), ("w_hh", model.spk.w_hh[l]), ("bias", model.spk.bias[l])):
            want = t[f"spk.{name}.{l}"].reshape(-1)
            assert torch.equal(_view(block, ptr, want.numel(), torch.float32), want), (name, l)
    assert torch.equal(_view(block, model.spk.lin_w, 256 * 256, torch.float32), t["spk.lin_w"].reshape(-1))
    assert torch.equal(_view(block, model.spk.lin_b, 256, torch.float32), t["spk.lin_b"])
    assert torch.equal(_view(block, model.tail.window, 16, torch.float32), t["tail.window"])
    assert torch.allclose(_view(block, model.tail.synth, 272, torch.float32), t["tail.synth"].reshape(-1), rtol=1e-6, atol=1e-9)
    assert torch.equal(tail_host[:16], t["tail.window"]) and torch.allclose(tail_host[16:], t["tail.synth"].reshape(-1), rtol=1e-6, atol=1e-9)
    print(f"opformat {opf}: worst fraction of elements off by one operand ulp {worst:.2e}")


def test_native_fold_reports_missing_and_misshapen_entries(sd):
    lib = capi.load()
    model = capi.Model()
    block = torch.zeros(int(lib.qvc_prepared_bytes(capi.OPF_TF32)), dtype=torch.uint8)
    bad = {k: v for k, v in sd.items() if k != "dec.ups.1.weight_g"}
    entries, _keep = capi.state_entries(bad)
    st = lib.qvc_fold_host(entries, len(entries), capi.OPF_TF32, capi.BACKEND_TCGEN05, block.data_ptr(), block.numel(), None, C.byref(model))
    assert st == -1 and b"dec.ups.1.weight_g" in lib.qvc_last_error()
    bad = dict(sd)
    bad["enc_p.pre.bias"] = torch.zeros(191)
    entries, _keep = capi.state_entries(bad)
    st = lib.qvc_fold_host(entries, len(entries), capi.OPF_TF32, capi.BACKEND_TCGEN05, block.data_ptr(), block.numel(), None, C.byref(model))
    assert st == -1 and b"enc_p.pre.bias" in lib.qvc_last_error()
    entries, _keep = capi.state_entries(sd)
    assert lib.qvc_fold_host(entries, len(entries), capi.OPF_F32, capi.BACKEND_TCGEN05, block.data_ptr(), block.numel(), None, C.byref(model)) == -1
    assert lib.qvc_fold_host(entries, len(entries), capi.OPF_TF32, capi.BACKEND_TCGEN05, block.data_ptr(), 4096, None, C.byref(model)) == -4
    assert lib.qvc_prepared_bytes(9) == 0
