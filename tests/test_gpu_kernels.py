"""Kernel-level parity on a B200, through the C ABI: series convolution (both back ends, every
epilogue and operand format), layout helpers, fused tail, persistent-RNN speaker encoder."""
import ctypes as C

import pytest
import torch

import synth
from gpu_util import conv1d, make_args, op_dtype, ref_conv, stream, to_op, tref
from oracle import qvc_oracle
from quickvc_official_b200 import capi, fold

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

BACKENDS = [("fma", capi.BACKEND_FMA, capi.OPF_F32), ("fma", capi.BACKEND_FMA, capi.OPF_TF32),
            ("fma", capi.BACKEND_FMA, capi.OPF_BF16),
            ("tc", capi.BACKEND_TCGEN05, capi.OPF_TF32), ("tc", capi.BACKEND_TCGEN05, capi.OPF_BF16),
            ("fma", capi.BACKEND_FMA, capi.OPF_F16), ("tc", capi.BACKEND_TCGEN05, capi.OPF_F16)]
IDS = ["fma-f32", "fma-tf32", "fma-bf16", "tc-tf32", "tc-bf16", "fma-f16", "tc-f16"]


def _tol(opf):
    # accumulation-order noise only: operands are pre-rounded on both sides
    return 2e-5 if opf != capi.OPF_BF16 else 1e-2


# (B, rows, cin, cout, k, dil): shapes of every layer family on the path plus ragged edge cases
GEOMS = [
    (2, 24, 256, 192, 1, 1),      # enc_p.pre
    (2, 37, 192, 192, 1, 1),      # flow pre/post (zero-embedded), ragged rows
    (1, 200, 192, 512, 7, 1),     # dec.conv_pre
    (2, 130, 256, 256, 3, 1),     # MRF-1 k3
    (1, 300, 256, 256, 11, 5),    # MRF-1 k11 dilation 5 (largest halo)
    (2, 515, 128, 128, 7, 3),     # MRF-2 k7 dilation 3
    (1, 129, 128, 80, 7, 1),      # conv_post (72 padded to 80)
    (3, 50, 512, 1280, 4, 1),     # ups.0 polyphase
    (1, 1, 192, 192, 5, 1),       # a single row
    (2, 5, 128, 128, 11, 5),      # series shorter than the filter span
    (2, 700, 256, 512, 5, 1),     # ups.1 polyphase: CTA-pair kernel, two channel groups, ragged last frame block
    (1, 257, 512, 1280, 4, 1),    # ups.0 polyphase: five channel groups, one frame past a pair tile
    (2, 129, 256, 256, 7, 5),     # one frame into the second CTA's half
]


@pytest.mark.parametrize("geom", GEOMS, ids=[f"B{g[0]}_R{g[1]}_{g[2]}to{g[3]}_k{g[4]}d{g[5]}" for g in GEOMS])
@pytest.mark.parametrize("be", BACKENDS, ids=IDS)
def test_conv_linear(geom, be):
    _, backend, opf = be
    B, rows, cin, cout, k, dil = geom
    g = torch.Generator(device="cpu").manual_seed(hash(geom) & 0xffff)
    x = to_op(torch.randn(B, rows, cin, generator=g), opf).to(DEV)
    w = to_op(torch.randn(cout, k, cin, generator=g) / (cin * k) ** 0.5, opf).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    res = torch.randn(B, rows, cout, generator=g).to(DEV)
    accin = torch.randn(B, rows, cout, generator=g).to(DEV)
    raw = torch.full((B, rows, cout), float("nan"), device=DEV)
    op = torch.zeros(B, rows, cout, device=DEV, dtype=op_dtype(opf))
    pad_left = (k - 1) * dil // 2
    conv1d(x, w, bias, k=k, dil=dil, pad_left=pad_left, out_rows=rows, opf=opf, backend=backend,
           segs=[dict(col0=0, ncols=cout, alpha=-1.0, beta=1.0 / 3, slope=0.1, res=res, accin=accin, raw=raw, op=op)])
    torch.cuda.synchronize()
    acc = ref_conv(x.cpu().float(), w.cpu().float(), k, dil, pad_left, rows)
    want = accin.cpu().double() + (1.0 / 3) * (-(acc + bias.cpu().double()) + res.cpu().double())
    scale = float(want.abs().max())
    assert float((raw.cpu().double() - want).abs().max()) < _tol(opf) * scale
    want_op = torch.where(want > 0, want, want * 0.1)
    err_op = float((op.cpu().double() - want_op).abs().max())
    assert err_op < (_tol(opf) + (2 ** -8 if opf == capi.OPF_BF16 else 2 ** -11 if opf in (capi.OPF_TF32, capi.OPF_F16) else 0)) * scale
    if opf == capi.OPF_TF32:        # the operand copy must sit on the TF32 grid
        assert int((op.view(torch.int32) & 0x1FFF).abs().sum()) == 0


@pytest.mark.parametrize("be", BACKENDS, ids=IDS)
@pytest.mark.parametrize("geom", [(2, 300, 128, 128, 3), (1, 700, 256, 256, 7), (13, 700, 256, 256, 7), (20, 1100, 512, 256, 3)],
                         ids=["128ch", "256ch-one-cta", "256ch-pairs-staged", "pairs-staged-two-rounds"])
def test_conv_residual_from_operand_copy(be, geom):
    """seg.res_op: the residual given as the operand-format copy of leaky_relu(r, 0.1)."""
    _, backend, opf = be
    B, rows, cin, cout, k = geom
    g = torch.Generator(device="cpu").manual_seed(3)
    x = to_op(torch.randn(B, rows, cin, generator=g), opf).to(DEV)
    w = to_op(torch.randn(cout, k, cin, generator=g) / (cin * k) ** 0.5, opf).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    r = torch.randn(B, rows, cout, generator=g)
    r_op = to_op(torch.where(r > 0, r, 0.1 * r), opf).to(DEV)            # what a producing epilogue stores
    raw = torch.full((B, rows, cout), float("nan"), device=DEV)
    conv1d(x, w, bias, k=k, dil=1, pad_left=(k - 1) // 2, out_rows=rows, opf=opf, backend=backend,
           segs=[dict(col0=0, ncols=cout, res_op=r_op, res_inv_slope=10.0, raw=raw)])
    torch.cuda.synchronize()
    acc = ref_conv(x.cpu().float(), w.cpu().float(), k, 1, (k - 1) // 2, rows)
    rq = r_op.cpu().double()
    want = acc + bias.cpu().double() + torch.where(rq > 0, rq, rq * 10.0)
    assert float((raw.cpu().double() - want).abs().max()) < _tol(opf) * float(want.abs().max())


@pytest.mark.parametrize("be", BACKENDS, ids=IDS)
def test_conv_two_segments_wn_res_skip(be):
    _, backend, opf = be
    B, rows, H = 2, 70, 192
    g = torch.Generator(device="cpu").manual_seed(5)
    acts = to_op(torch.randn(B, rows, H, generator=g), opf).to(DEV)
    w = to_op(torch.randn(2 * H, 1, H, generator=g) / H ** 0.5, opf).to(DEV)
    bias = torch.randn(2 * H, generator=g).to(DEV)
    x = torch.randn(B, rows, H, generator=g).to(DEV)
    skip = torch.randn(B, rows, H, generator=g).to(DEV)
    x0, skip0 = x.clone(), skip.clone()
    xo = torch.zeros(B, rows, H, device=DEV, dtype=op_dtype(opf))
    conv1d(acts, w, bias, k=1, dil=1, pad_left=0, out_rows=rows, opf=opf, backend=backend,
           segs=[dict(col0=0, ncols=H, res=x, raw=x, op=xo), dict(col0=H, ncols=H, accin=skip, raw=skip)])
    torch.cuda.synchronize()
    acc = ref_conv(acts.cpu().float(), w.cpu().float(), 1, 1, 0, rows) + bias.cpu().double()
    assert float((x.cpu().double() - (x0.cpu().double() + acc[..., :H])).abs().max()) < _tol(opf) * 4
    assert float((skip.cpu().double() - (skip0.cpu().double() + acc[..., H:])).abs().max()) < _tol(opf) * 4


@pytest.mark.parametrize("opf", [capi.OPF_TF32, capi.OPF_BF16, capi.OPF_F16], ids=["tf32", "bf16", "f16"])
@pytest.mark.parametrize("last", [False, True], ids=["res_skip", "skip_only"])
def test_fused_wn_layer_equals_the_two_convolutions(opf, last, monkeypatch):
    """qvc_wn_layer (one CTA-pair kernel, activations kept in shared memory) against qvc_conv1d(in_layer) +
    qvc_conv1d(res_skip) on the same operands: same arithmetic, so the results must agree to accumulation noise."""
    monkeypatch.setenv("QVC_WN_FUSED", "1")          # off by default since the frames-on-rows kernel serves the WN stacks
    lib = capi.load()
    B, rows, H, k = 9, 700, 192, 5                  # 9 x 6 pair tiles; the last frame block is ragged (700 = 5*128 + 60)
    g = torch.Generator(device="cpu").manual_seed(11)
    x = to_op(torch.randn(B, rows, H, generator=g), opf).to(DEV)
    w_in = to_op(torch.randn(2 * H, k, H, generator=g) / (H * k) ** 0.5, opf).to(DEV)
    gbias = torch.randn(B, 2 * H, generator=g).to(DEV)
    rs_out = H if last else 2 * H
    w_rs = to_op(torch.randn(rs_out, 1, H, generator=g) / H ** 0.5, opf).to(DEV)
    b_rs = torch.randn(rs_out, generator=g).to(DEV)
    xr0 = torch.randn(B, rows, H, generator=g).to(DEV)
    sk0 = torch.randn(B, rows, H, generator=g).to(DEV)
    be = capi.BACKEND_TCGEN05

    def run(fused):
        xr, sk = xr0.clone(), sk0.clone()
        acts = torch.zeros(B, rows, H, device=DEV, dtype=op_dtype(opf))
        xo = torch.zeros(B, rows, H, device=DEV, dtype=op_dtype(opf))
        a = make_args(x, w_in, gbias, k=k, dil=1, pad_left=2, out_rows=rows, opf=opf, backend=be, epilogue=capi.EPI_GATE,
                      segs=[dict(col0=0, ncols=H, op=acts)], bias_bstride=2 * H)
        if last:
            segs = [dict(col0=0, ncols=H, accin=sk, op=xo)]
        else:
            segs = [dict(col0=0, ncols=H, res=xr, raw=xr, op=xo), dict(col0=H, ncols=H, accin=sk, raw=sk)]
        r = make_args(acts, w_rs, b_rs, k=1, dil=1, pad_left=0, out_rows=rows, opf=opf, backend=be, segs=segs)
        if fused:
            capi.check(lib.qvc_wn_layer(C.byref(a), C.byref(r), stream()), "qvc_wn_layer")
        else:
            capi.check(lib.qvc_conv1d(C.byref(a), stream()), "qvc_conv1d")
            capi.check(lib.qvc_conv1d(C.byref(r), stream()), "qvc_conv1d")
        torch.cuda.synchronize()
        return xr.cpu(), sk.cpu(), xo.float().cpu()

    want, got = run(False), run(True)
    tol = 2e-5 if opf == capi.OPF_TF32 else 1e-2
    for name, w, gt in zip(("x", "skip", "x operand"), want, got):
        scale = float(w.abs().max()) + 1e-6
        assert float((w - gt).abs().max()) < tol * scale, name


TC_OPFS = [capi.OPF_TF32, capi.OPF_BF16, capi.OPF_F16]
TC_IDS = ["tf32", "bf16", "f16"]


def _force_rows(monkeypatch):
    monkeypatch.setenv("QVC_TC_ROWS", "7")           # every eligible layer, not only the 192 / 384-column ones
    monkeypatch.setenv("QVC_TC_2CTA_FORCE", "1")     # ... however few tiles it has


@pytest.mark.parametrize("opf", TC_OPFS, ids=TC_IDS)
@pytest.mark.parametrize("per_utt_bias", [False, True])
@pytest.mark.parametrize("rows", [129, 300, 700])
def test_rows_kernel_gate(opf, per_utt_bias, rows, monkeypatch):
    """WN in_layer + gate on the frames-on-rows pair kernel (conv_tcr.cu): two 96-channel pieces per 256-frame tile."""
    _force_rows(monkeypatch)
    B, H, k = 3, 192, 5
    g = torch.Generator(device="cpu").manual_seed(7 + rows)
    x = to_op(torch.randn(B, rows, H, generator=g), opf).to(DEV)
    w = to_op(torch.randn(2 * H, k, H, generator=g) / (H * k) ** 0.5, opf).to(DEV)
    bias = torch.randn(B if per_utt_bias else 1, 2 * H, generator=g).to(DEV)
    op = torch.zeros(B, rows, H, device=DEV, dtype=op_dtype(opf))
    raw = torch.zeros(B, rows, H, device=DEV)
    conv1d(x, w, bias, k=k, dil=1, pad_left=2, out_rows=rows, opf=opf, backend=capi.BACKEND_TCGEN05, epilogue=capi.EPI_GATE,
           segs=[dict(col0=0, ncols=H, op=op, raw=raw)], bias_bstride=2 * H if per_utt_bias else 0)
    assert capi.last_kernel() == "conv_tcr_kernel"
    torch.cuda.synchronize()
    a = ref_conv(x.cpu().float(), w.cpu().float(), k, 1, 2, rows) + bias.cpu().double().reshape(-1, 1, 2 * H)
    want = torch.tanh(a[..., :H]) * torch.sigmoid(a[..., H:])
    assert float((raw.cpu().double() - want).abs().max()) < max(_tol(opf), 1e-5)
    assert float((op.cpu().double() - want).abs().max()) < max(_tol(opf), 1e-5) + 2 ** -8


@pytest.mark.parametrize("opf", TC_OPFS, ids=TC_IDS)
@pytest.mark.parametrize("last", [False, True], ids=["res_skip", "skip_only"])
@pytest.mark.parametrize("rows", [130, 700])
@pytest.mark.parametrize("lean", [True, False], ids=["lean", "generic"])
def test_rows_kernel_res_skip(opf, last, rows, lean, monkeypatch):
    """WN res_skip 1x1 with its two-segment epilogue (x += res in place, skip += skip) on the frames-on-rows kernel, on
    its lean epilogue instance (one segment per column group) and on the generic one."""
    _force_rows(monkeypatch)
    monkeypatch.setenv("QVC_TCR_LEAN", "1" if lean else "0")
    B, H = 2, 192
    g = torch.Generator(device="cpu").manual_seed(5 + rows)
    acts = to_op(torch.randn(B, rows, H, generator=g), opf).to(DEV)
    cout = H if last else 2 * H
    w = to_op(torch.randn(cout, 1, H, generator=g) / H ** 0.5, opf).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    x = torch.randn(B, rows, H, generator=g).to(DEV)
    skip = torch.randn(B, rows, H, generator=g).to(DEV)
    x0, skip0 = x.clone(), skip.clone()
    xo = torch.zeros(B, rows, H, device=DEV, dtype=op_dtype(opf))
    if last:
        segs = [dict(col0=0, ncols=H, accin=skip, op=xo)]
    else:
        segs = [dict(col0=0, ncols=H, res=x, raw=x, op=xo), dict(col0=H, ncols=H, accin=skip, raw=skip)]
    conv1d(acts, w, bias, k=1, dil=1, pad_left=0, out_rows=rows, opf=opf, backend=capi.BACKEND_TCGEN05, segs=segs)
    assert capi.last_kernel() == "conv_tcr_kernel"
    torch.cuda.synchronize()
    acc = ref_conv(acts.cpu().float(), w.cpu().float(), 1, 1, 0, rows) + bias.cpu().double()
    tol = _tol(opf) * 4
    rnd = 2 ** -8 if opf == capi.OPF_BF16 else 2 ** -11
    if last:
        want = skip0.cpu().double() + acc
        assert float((xo.cpu().double() - want).abs().max()) < tol + rnd * float(want.abs().max())
        assert torch.equal(skip, skip0) and torch.equal(x, x0)
    else:
        want_x = x0.cpu().double() + acc[..., :H]
        assert float((x.cpu().double() - want_x).abs().max()) < tol
        assert float((skip.cpu().double() - (skip0.cpu().double() + acc[..., H:])).abs().max()) < tol
        assert float((xo.cpu().double() - want_x).abs().max()) < tol + rnd * float(want_x.abs().max())


# (B, rows, cin, cout, k, dil): piece widths 192 (enc_p.pre, flow pre / post), 256, 128, 2 x 256, 5 x 256
ROW_GEOMS = [(3, 300, 256, 192, 1, 1), (2, 515, 192, 192, 1, 1), (2, 130, 256, 256, 3, 1), (2, 515, 128, 128, 7, 3),
             (1, 300, 256, 256, 11, 5), (2, 700, 256, 512, 5, 1), (1, 257, 512, 1280, 4, 1)]


@pytest.mark.parametrize("geom", ROW_GEOMS, ids=[f"B{g[0]}_R{g[1]}_{g[2]}to{g[3]}_k{g[4]}d{g[5]}" for g in ROW_GEOMS])
@pytest.mark.parametrize("opf", TC_OPFS, ids=TC_IDS)
def test_rows_kernel_linear(geom, opf, monkeypatch):
    """General LINEAR epilogue (alpha, beta, slope, residual, accumulate-into, raw + operand outputs) on the
    frames-on-rows kernel, for every piece width it tiles outputs with."""
    _force_rows(monkeypatch)
    B, rows, cin, cout, k, dil = geom
    g = torch.Generator(device="cpu").manual_seed(hash(geom) & 0xffff)
    x = to_op(torch.randn(B, rows, cin, generator=g), opf).to(DEV)
    w = to_op(torch.randn(cout, k, cin, generator=g) / (cin * k) ** 0.5, opf).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    res = torch.randn(B, rows, cout, generator=g).to(DEV)
    accin = torch.randn(B, rows, cout, generator=g).to(DEV)
    raw = torch.full((B, rows, cout), float("nan"), device=DEV)
    op = torch.zeros(B, rows, cout, device=DEV, dtype=op_dtype(opf))
    pad_left = (k - 1) * dil // 2
    conv1d(x, w, bias, k=k, dil=dil, pad_left=pad_left, out_rows=rows, opf=opf, backend=capi.BACKEND_TCGEN05,
           segs=[dict(col0=0, ncols=cout, alpha=-1.0, beta=1.0 / 3, slope=0.1, res=res, accin=accin, raw=raw, op=op)])
    assert capi.last_kernel() == "conv_tcr_kernel"
    torch.cuda.synchronize()
    acc = ref_conv(x.cpu().float(), w.cpu().float(), k, dil, pad_left, rows)
    want = accin.cpu().double() + (1.0 / 3) * (-(acc + bias.cpu().double()) + res.cpu().double())
    scale = float(want.abs().max())
    assert float((raw.cpu().double() - want).abs().max()) < _tol(opf) * scale
    want_op = torch.where(want > 0, want, want * 0.1)
    assert float((op.cpu().double() - want_op).abs().max()) < (_tol(opf) + (2 ** -8 if opf == capi.OPF_BF16 else 2 ** -11)) * scale


@pytest.mark.parametrize("opf", TC_OPFS, ids=TC_IDS)
@pytest.mark.parametrize("geom", [(2, 300, 128, 128, 3), (3, 700, 128, 128, 7), (1, 515, 256, 256, 11)],
                         ids=["128ch-k3", "128ch-k7", "256ch-k11"])
def test_rows_kernel_residual_from_operand_copy(opf, geom, monkeypatch):
    """seg.res_op on the frames-on-rows kernel (the c2 layers of MRF-2 in the 16-bit modes): the residual arrives as the
    operand-format copy of leaky_relu(r, 0.1); tap-grouped filter stages (3, 4 + 3 and 2 x 6 taps per TMA operation)."""
    _force_rows(monkeypatch)
    B, rows, cin, cout, k = geom
    g = torch.Generator(device="cpu").manual_seed(3)
    x = to_op(torch.randn(B, rows, cin, generator=g), opf).to(DEV)
    w = to_op(torch.randn(cout, k, cin, generator=g) / (cin * k) ** 0.5, opf).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    r = torch.randn(B, rows, cout, generator=g)
    r_op = to_op(torch.where(r > 0, r, 0.1 * r), opf).to(DEV)
    raw = torch.full((B, rows, cout), float("nan"), device=DEV)
    op = torch.zeros(B, rows, cout, device=DEV, dtype=op_dtype(opf))
    conv1d(x, w, bias, k=k, dil=1, pad_left=(k - 1) // 2, out_rows=rows, opf=opf, backend=capi.BACKEND_TCGEN05,
           segs=[dict(col0=0, ncols=cout, slope=0.1, res_op=r_op, res_inv_slope=10.0, raw=raw, op=op)])
    assert capi.last_kernel() == "conv_tcr_kernel"
    torch.cuda.synchronize()
    acc = ref_conv(x.cpu().float(), w.cpu().float(), k, 1, (k - 1) // 2, rows)
    rq = r_op.cpu().double()
    want = acc + bias.cpu().double() + torch.where(rq > 0, rq, rq * 10.0)
    scale = float(want.abs().max())
    assert float((raw.cpu().double() - want).abs().max()) < _tol(opf) * scale
    want_op = torch.where(want > 0, want, want * 0.1)
    assert float((op.cpu().double() - want_op).abs().max()) < (_tol(opf) + (2 ** -8 if opf == capi.OPF_BF16 else 2 ** -11)) * scale


@pytest.mark.parametrize("be", BACKENDS, ids=IDS)
@pytest.mark.parametrize("per_utt_bias", [False, True])
@pytest.mark.parametrize("rows", [45, 300], ids=["R45", "R300"])      # R300: the CTA-pair (cta_group::2) gate kernel
def test_conv_gate(be, per_utt_bias, rows):
    _, backend, opf = be
    B, H, k = 3, 192, 5
    g = torch.Generator(device="cpu").manual_seed(7)
    x = to_op(torch.randn(B, rows, H, generator=g), opf).to(DEV)
    w = to_op(torch.randn(2 * H, k, H, generator=g) / (H * k) ** 0.5, opf).to(DEV)
    bias = torch.randn(B if per_utt_bias else 1, 2 * H, generator=g).to(DEV)
    op = torch.zeros(B, rows, H, device=DEV, dtype=op_dtype(opf))
    raw = torch.zeros(B, rows, H, device=DEV)
    conv1d(x, w, bias, k=k, dil=1, pad_left=2, out_rows=rows, opf=opf, backend=backend, epilogue=capi.EPI_GATE,
           segs=[dict(col0=0, ncols=H, op=op, raw=raw)], bias_bstride=2 * H if per_utt_bias else 0)
    torch.cuda.synchronize()
    a = ref_conv(x.cpu().float(), w.cpu().float(), k, 1, 2, rows) + bias.cpu().double().reshape(-1, 1, 2 * H)
    want = torch.tanh(a[..., :H]) * torch.sigmoid(a[..., H:])
    assert float((raw.cpu().double() - want).abs().max()) < max(_tol(opf), 1e-5)
    assert float((op.cpu().double() - want).abs().max()) < max(_tol(opf), 1e-5) + 2 ** -8


@pytest.mark.parametrize("be", BACKENDS, ids=IDS)
def test_conv_sample(be):
    _, backend, opf = be
    B, rows, H = 2, 33, 192
    g = torch.Generator(device="cpu").manual_seed(9)
    x = to_op(torch.randn(B, rows, H, generator=g), opf).to(DEV)
    w = to_op(torch.randn(2 * H, 1, H, generator=g) / H ** 0.5, opf).to(DEV)
    bias = (0.1 * torch.randn(2 * H, generator=g)).to(DEV)
    noise = torch.randn(B, rows, H, generator=g).to(DEV)
    z = torch.zeros(B, rows, H, device=DEV)
    m = torch.zeros(B, rows, H, device=DEV)
    lg = torch.zeros(B, rows, H, device=DEV)
    zo = torch.zeros(B, rows, H, device=DEV, dtype=op_dtype(opf))
    conv1d(x, w, bias, k=1, dil=1, pad_left=0, out_rows=rows, opf=opf, backend=backend, epilogue=capi.EPI_SAMPLE,
           segs=[dict(col0=0, ncols=H, raw=z, op=zo)], noise=noise, aux0=m, aux1=lg)
    torch.cuda.synchronize()
    a = ref_conv(x.cpu().float(), w.cpu().float(), 1, 1, 0, rows) + bias.cpu().double()
    want = a[..., :H] + noise.cpu().double() * torch.exp(a[..., H:])
    tol = max(_tol(opf), 1e-5) * float(want.abs().max())
    assert float((z.cpu().double() - want).abs().max()) < tol
    assert float((m.cpu().double() - a[..., :H]).abs().max()) < tol
    assert float((lg.cpu().double() - a[..., H:]).abs().max()) < tol


def test_layout_roundtrip():
    lib = capi.load()
    x = torch.randn(3, 80, 77, device=DEV)
    sm = torch.empty(3, 77, 80, device=DEV)
    capi.check(lib.qvc_to_series_major(x.data_ptr(), sm.data_ptr(), 3, 80, 77, capi.OPF_F32, stream()), "to")
    assert torch.equal(sm, x.transpose(1, 2).contiguous())
    back = torch.empty_like(x)
    capi.check(lib.qvc_from_series_major(sm.data_ptr(), 80, back.data_ptr(), 3, 80, 77, stream()), "from")
    assert torch.equal(back, x)
    t32 = torch.empty(3, 77, 80, device=DEV)
    capi.check(lib.qvc_to_series_major(x.data_ptr(), t32.data_ptr(), 3, 80, 77, capi.OPF_TF32, stream()), "to")
    assert torch.equal(t32.cpu(), fold.round_tf32(x.transpose(1, 2).contiguous().cpu()))
    b16 = torch.empty(3, 77, 80, device=DEV, dtype=torch.bfloat16)
    capi.check(lib.qvc_to_series_major(x.data_ptr(), b16.data_ptr(), 3, 80, 77, capi.OPF_BF16, stream()), "to")
    assert torch.equal(b16, x.transpose(1, 2).contiguous().to(torch.bfloat16))


@pytest.mark.parametrize("frames,batch", [(2, 1), (17, 3), (481, 2), (1001, 1)])
def test_tail_matches_oracle(frames, batch, sd):
    lib = capi.load()
    g = torch.Generator(device="cpu").manual_seed(frames)
    post = 0.5 * torch.randn(batch, 72, frames, generator=g)
    f = fold.fold_state_dict({k: v for k, v in sd.items() if not k.startswith("enc_q.")}, capi.OPF_F32)
    tw = capi.TailWeights(f.tensors["tail.window"].to(DEV).data_ptr(), 0)
    win, syn = f.tensors["tail.window"].to(DEV), f.tensors["tail.synth"].to(DEV)
    tw = capi.TailWeights(win.data_ptr(), syn.data_ptr())
    post_sm = post.transpose(1, 2).contiguous().to(DEV)
    wave = torch.full((batch, 1, 16 * (frames - 1)), float("nan"), device=DEV)
    ymb = torch.full((batch, 4, 4 * (frames - 1)), float("nan"), device=DEV)
    capi.check(lib.qvc_tail(C.byref(tw), post_sm.data_ptr(), 72, batch, frames, None, 0, wave.data_ptr(), ymb.data_ptr(),
                            stream()), "qvc_tail")
    torch.cuda.synchronize()
    # oracle: the decoder code after subband_conv_post (models.py:390-406)
    p = qvc_oracle._P(sd, torch.float64)
    x = post.double().reshape(batch, 4, 18, frames)
    y = qvc_oracle.istft_closed_form(x[:, :, :9].reshape(batch * 4, 9, frames), x[:, :, 9:].reshape(batch * 4, 9, frames),
                                     p.raw("dec.stft.window")).reshape(batch, 4, -1)
    import torch.nn.functional as F
    up = F.conv_transpose1d(y, p.raw("dec.updown_filter") * 4, stride=4)
    want = F.conv1d(up, p.weight("dec.multistream_conv_post"), padding=31)
    assert synth.max_abs(ymb, y) < 2e-5 * float(y.abs().max())
    assert synth.max_abs(wave, want) < 2e-5 * float(want.abs().max())
    # same kernel with the coefficients passed as kernel parameters (host copies given): identical arithmetic
    tw2 = capi.TailWeights(win.data_ptr(), syn.data_ptr(), f.tensors["tail.window_host"].data_ptr(),
                           f.tensors["tail.synth_host"].data_ptr())
    wave2 = torch.full_like(wave, float("nan"))
    capi.check(lib.qvc_tail(C.byref(tw2), post_sm.data_ptr(), 72, batch, frames, None, 0, wave2.data_ptr(), None, stream()), "qvc_tail")
    torch.cuda.synchronize()
    assert torch.equal(wave2, wave)


@pytest.mark.parametrize("opf", TC_OPFS, ids=TC_IDS)
@pytest.mark.parametrize("frames,batch,ragged", [(2, 1, False), (17, 3, False), (122, 2, False), (123, 1, False), (481, 2, True),
                                                 (1001, 3, True), (10001, 2, False)])
def test_fused_post_net_and_tail_equals_the_two_launches(opf, frames, batch, ragged, sd):
    """qvc_post_tail (subband_conv_post as a tcgen05 GEMM whose epilogue is the tail) against qvc_conv1d + qvc_tail on the
    same operands: tile edges (121 frames per CTA, 242 per pair), the overlap-add / FIR halo recomputed per tile, ragged
    lengths, and the optional post-net and sub-band taps."""
    lib = capi.load()
    be = capi.BACKEND_TCGEN05
    g = torch.Generator(device="cpu").manual_seed(frames)
    x = to_op(0.5 * torch.randn(batch, frames, 128, generator=g), opf).to(DEV)
    w = to_op(torch.randn(80, 7, 128, generator=g) / (128 * 7) ** 0.5, opf)
    w[72:] = 0
    w = w.to(DEV)
    bias = (0.3 * torch.randn(80, generator=g))
    bias[72:] = 0
    bias = bias.to(DEV)
    f = fold.fold_state_dict({k: v for k, v in sd.items() if not k.startswith("enc_q.")}, capi.OPF_F32)
    win, syn = f.tensors["tail.window"].to(DEV), f.tensors["tail.synth"].to(DEV)
    tw = capi.TailWeights(win.data_ptr(), syn.data_ptr(), f.tensors["tail.window_host"].data_ptr(),
                          f.tensors["tail.synth_host"].data_ptr())
    fpu = 20
    lens = None
    if ragged:                                     # frames = 20 * units + 1
        units = (frames - 1) // fpu
        lens = torch.tensor([max(1, units - 7 * i) for i in range(batch)], dtype=torch.int32, device=DEV)
    lp = lens.data_ptr() if lens is not None else None
    # two launches: the convolution writes the 72-channel tensor, the tail reads it back
    post = torch.zeros(batch, frames, 72, device=DEV)
    a = make_args(x, w, bias, k=7, dil=1, pad_left=3, out_rows=frames, opf=opf, backend=be, segs=[dict(col0=0, ncols=72, raw=post)])
    capi.check(lib.qvc_conv1d(C.byref(a), stream()), "qvc_conv1d")
    wave0 = torch.full((batch, 1, 16 * (frames - 1)), float("nan"), device=DEV)
    ymb0 = torch.full((batch, 4, 4 * (frames - 1)), float("nan"), device=DEV)
    capi.check(lib.qvc_tail(C.byref(tw), post.data_ptr(), 72, batch, frames, lp, fpu, wave0.data_ptr(), ymb0.data_ptr(), stream()), "qvc_tail")
    # one launch, with both taps
    post1 = torch.full((batch, frames, 72), float("nan"), device=DEV)
    a1 = make_args(x, w, bias, k=7, dil=1, pad_left=3, out_rows=frames, opf=opf, backend=be, segs=[dict(col0=0, ncols=72, raw=post1)])
    wave1 = torch.full_like(wave0, float("nan"))
    ymb1 = torch.full_like(ymb0, float("nan"))
    capi.check(lib.qvc_post_tail(C.byref(a1), C.byref(tw), lp, fpu, wave1.data_ptr(), ymb1.data_ptr(), stream()), "qvc_post_tail")
    assert capi.last_kernel() == "post_tail_kernel"
    # ... and without them
    a2 = make_args(x, w, bias, k=7, dil=1, pad_left=3, out_rows=frames, opf=opf, backend=be, segs=[dict(col0=0, ncols=72)])
    wave2 = torch.full_like(wave0, float("nan"))
    capi.check(lib.qvc_post_tail(C.byref(a2), C.byref(tw), lp, fpu, wave2.data_ptr(), None, stream()), "qvc_post_tail")
    torch.cuda.synchronize()
    tol = 2e-5 if opf != capi.OPF_BF16 else 2e-3         # accumulation order of the GEMM only: operands are pre-rounded
    scale = float(wave0.abs().max()) + 1e-9
    assert not torch.isnan(wave1).any() and not torch.isnan(ymb1).any()
    assert float((wave1 - wave0).abs().max()) < tol * scale
    assert float((ymb1 - ymb0).abs().max()) < tol * (float(ymb0.abs().max()) + 1e-9)
    assert torch.equal(wave2, wave1)
    if lens is None:
        assert float((post1 - post).abs().max()) < tol * float(post.abs().max())
    else:
        for b in range(batch):
            n = int(lens[b]) * fpu + 1
            assert float((post1[b, :n] - post[b, :n]).abs().max()) < tol * float(post.abs().max())
            assert float(wave1[b, 0, 16 * (n - 1):].abs().max() if n < frames else 0.0) == 0.0


@pytest.mark.parametrize("bm,tm", [(1, 129), (1, 250), (1, 500), (1, 1500), (3, 100), (1, 128), (9, 7)])
@pytest.mark.parametrize("spc", [None, 1, 4, 8])
def test_speaker_encoder_matches_oracle(bm, tm, spc, sd, monkeypatch):
    # spc = sequences (windows) per LSTM cluster: one template instance of the recurrent kernel each
    if spc is not None:
        monkeypatch.setenv("QVC_SPK_SPC", str(spc))
    lib = capi.load()
    f = fold.fold_state_dict({k: v for k, v in sd.items() if not k.startswith("enc_q.")}, capi.OPF_F32)
    t = {k: v.to(DEV) for k, v in f.tensors.items() if k.startswith("spk.")}
    sw = capi.SpkWeights()
    for l in range(3):
        sw.w_ih[l], sw.w_hh[l], sw.bias[l] = t[f"spk.w_ih.{l}"].data_ptr(), t[f"spk.w_hh.{l}"].data_ptr(), t[f"spk.bias.{l}"].data_ptr()
    sw.lin_w, sw.lin_b = t["spk.lin_w"].data_ptr(), t["spk.lin_b"].data_ptr()
    _, mel, _ = synth.synthetic_inputs(1, 1, bm, tm, 11)
    want = qvc_oracle.embed_utterance(qvc_oracle._P(sd, torch.float64), mel.double())
    n_embed = want.shape[0]
    g = torch.full((n_embed, 256), float("nan"), device=DEV)
    nbytes = int(lib.qvc_spk_workspace_bytes(bm, tm))
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=DEV)
    base = (ws.data_ptr() + 255) & ~255
    mel_d = mel.to(DEV)
    capi.check(lib.qvc_spk_embed(C.byref(sw), mel_d.data_ptr(), bm, tm, g.data_ptr(), base, nbytes, stream()), "spk")
    torch.cuda.synchronize()
    assert synth.rel_l2(g, want) < 1e-4, synth.rel_l2(g, want)


def test_error_reporting():
    lib = capi.load()
    a = capi.ConvArgs()
    assert lib.qvc_conv1d(C.byref(a), stream()) == -1
    assert b"null" in lib.qvc_last_error()
    sw = capi.SpkWeights()
    mel = torch.zeros(2, 80, 200, device=DEV)
    g = torch.zeros(1, 256, device=DEV)
    ws = torch.zeros(1 << 20, dtype=torch.uint8, device=DEV)
    assert lib.qvc_spk_embed(C.byref(sw), mel.data_ptr(), 2, 200, g.data_ptr(), ws.data_ptr(), ws.numel(), stream()) == -1
    assert b"batch 1" in lib.qvc_last_error()


@pytest.mark.parametrize("opf", [capi.OPF_TF32, capi.OPF_BF16, capi.OPF_F16], ids=["tf32", "bf16", "f16"])
@pytest.mark.parametrize("k,batch,rows", [(3, 2, 300), (7, 1, 1030), (11, 40, 640)], ids=["k3", "k7", "k11-pairs"])
def test_frame_paired_convolution_with_tap_hints(opf, k, batch, rows):
    """qvc_model.paired / qvc_conv_args.tap_split: a 128 -> 128 dilation-1 layer run as the frame-paired 256 -> 256
    layer on the [rows/2][256] view equals the plain layer, and the structured-zero hint changes no bit."""
    g = torch.Generator().manual_seed(k)
    C, pad = 128, (k - 1) // 2
    x = to_op(torch.randn(batch, rows, C, generator=g), opf).to(DEV)
    w32 = torch.randn(C, k, C, generator=g) / (C * k) ** 0.5
    w = to_op(w32, opf).to(DEV)
    bias = torch.randn(C, generator=g).to(DEV)
    res = torch.randn(batch, rows, C, generator=g).to(DEV)
    kw = dict(opf=opf, backend=capi.BACKEND_TCGEN05)

    def run(xv, wv, bv, kk, padl, nrows, width, taps=None):
        raw = torch.full((batch, nrows, width), float("nan"), device=DEV)
        op = torch.empty(batch, nrows, width, device=DEV, dtype=op_dtype(opf))
        segs = [dict(col0=0, ncols=width, slope=0.1, res=res.view(batch, nrows, width), raw=raw, op=op)]
        conv1d(xv, wv, bv, k=kk, dil=1, pad_left=padl, out_rows=nrows, segs=segs, taps=taps, **kw)
        torch.cuda.synchronize()
        return raw.view(batch, rows, C), op.view(batch, rows, C)

    plain_raw, plain_op = run(x, w, bias, k, pad, rows, C)
    wp32, pad_p = fold.frame_pair_filter(w.float().cpu(), pad)          # from the already rounded operands: exact
    wp = to_op(wp32, opf).to(DEV)
    kp = wp.shape[1]
    lo = [[0, 0], [0, 0]]
    hi = [[0, 0], [0, 0]]
    for p in range(2):
        for q in range(2):
            nz = [a for a in range(kp) if float(wp32[p * C:(p + 1) * C, a, q * C:(q + 1) * C].abs().sum()) > 0]
            lo[p][q], hi[p][q] = nz[0], nz[-1]
    xp, bp = x.view(batch, rows // 2, 2 * C), torch.cat([bias, bias])
    hint_raw, hint_op = run(xp, wp, bp, kp, pad_p, rows // 2, 2 * C, taps=(C, lo, hi))
    full_raw, full_op = run(xp, wp, bp, kp, pad_p, rows // 2, 2 * C)
    assert torch.equal(hint_raw, full_raw) and torch.equal(hint_op, full_op)        # skipped blocks are zeros
    want = ref_conv(x, w, k, 1, pad, rows) + bias.double() + res.double()
    assert float((hint_raw.double() - want).abs().max()) < _tol(opf)
    assert float((hint_raw - plain_raw).abs().max()) < _tol(opf)                    # summation order only


# ------------------------------------------------------------------------------------------------
# qvc_conv1d_sum: a sum of convolutions accumulated in tensor memory with one multi-residual epilogue
# (the mean of the three ResBlocks of an MRF stage, models.py:378-384)
# ------------------------------------------------------------------------------------------------
def _sum_case(opf, B, rows, ch, ks, res_mode, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    srcs = []
    for k in ks:
        x = to_op(torch.randn(B, rows, ch, generator=g), opf).to(DEV)
        w = to_op(torch.randn(ch, k, ch, generator=g) / (ch * k) ** 0.5, opf).to(DEV)
        bias = torch.randn(ch, generator=g).to(DEV)
        r = torch.randn(B, rows, ch, generator=g)
        r_op = to_op(torch.where(r > 0, r, 0.1 * r), opf).to(DEV)
        srcs.append(dict(x=x, w=w, bias=bias, k=k, res=r.to(DEV), res_op=r_op))
    return srcs


def _run_sum(srcs, opf, rows, ch, res_mode, raw, op):
    lib = capi.load()
    args = []
    for i, s in enumerate(srcs):
        seg = dict(col0=0, ncols=ch)
        if res_mode == "res":
            seg["res"] = s["res"]
        elif res_mode == "res_op":
            seg["res_op"], seg["res_inv_slope"] = s["res_op"], 10.0
        if i == 0:
            seg.update(beta=1.0 / 3, slope=0.01, raw=raw, op=op)
        args.append(make_args(s["x"], s["w"], s["bias"], k=s["k"], dil=1, pad_left=(s["k"] - 1) // 2, out_rows=rows, opf=opf,
                              backend=capi.BACKEND_TCGEN05, segs=[seg]))
    arr = (C.POINTER(capi.ConvArgs) * len(args))(*[C.pointer(a) for a in args])
    capi.check(lib.qvc_conv1d_sum(arr, len(args), stream()), "qvc_conv1d_sum")
    torch.cuda.synchronize()


@pytest.mark.parametrize("opf", [capi.OPF_TF32, capi.OPF_BF16, capi.OPF_F16], ids=["tf32", "bf16", "f16"])
@pytest.mark.parametrize("res_mode", ["res", "res_op", "none"])
@pytest.mark.parametrize("geom", [(2, 300, 128, (3, 7, 11)), (9, 700, 256, (3, 7, 11)), (1, 37, 256, (11, 3)), (3, 129, 128, (7,))],
                         ids=["128ch", "256ch-pairs", "two-sources-short", "one-source"])
def test_conv_sum_matches_reference(opf, res_mode, geom, monkeypatch):
    B, rows, ch, ks = geom
    srcs = _sum_case(opf, B, rows, ch, ks, res_mode, 17)
    want = torch.zeros(B, rows, ch, dtype=torch.float64)
    for s in srcs:
        want += ref_conv(s["x"].cpu().float(), s["w"].cpu().float(), s["k"], 1, (s["k"] - 1) // 2, rows) + s["bias"].cpu().double()
        if res_mode == "res":
            want += s["res"].cpu().double()
        elif res_mode == "res_op":
            rq = s["res_op"].cpu().double()
            want += torch.where(rq > 0, rq, rq * 10.0)
    want = want / 3
    results = []
    for force in ("0", "1"):                       # one-CTA kernel, then (where the shape allows) the CTA-pair kernel
        monkeypatch.setenv("QVC_TC_2CTA_FORCE", force)
        monkeypatch.setenv("QVC_TC_2CTA", force)
        raw = torch.full((B, rows, ch), float("nan"), device=DEV)
        op = torch.zeros(B, rows, ch, device=DEV, dtype=op_dtype(opf))
        _run_sum(srcs, opf, rows, ch, res_mode, raw, op)
        scale = float(want.abs().max())
        assert float((raw.cpu().double() - want).abs().max()) < _tol(opf) * scale
        want_op = torch.where(want > 0, want, want * 0.01)
        assert float((op.cpu().double() - want_op).abs().max()) < (_tol(opf) + (2 ** -8 if opf == capi.OPF_BF16 else 2 ** -11)) * scale
        results.append((raw.clone(), op.clone()))
    # the two kernels accumulate the same K blocks in the same order: bit-identical (an utterance's samples must not
    # depend on which kernel its batch size selects)
    assert torch.equal(results[0][0], results[1][0]) and torch.equal(results[0][1], results[1][1])


def test_conv_sum_rejects_mismatched_sources():
    opf = capi.OPF_TF32
    a = _sum_case(opf, 1, 64, 128, (3, 7), "res", 1)
    b = _sum_case(opf, 1, 64, 256, (3,), "res", 2)
    raw = torch.zeros(1, 64, 128, device=DEV)
    with pytest.raises(capi.QvcError):
        _run_sum([a[0], b[0]], opf, 64, 128, "res", raw, None)
    with pytest.raises(capi.QvcError):               # mixed residual kinds
        lib = capi.load()
        a0 = make_args(a[0]["x"], a[0]["w"], a[0]["bias"], k=3, dil=1, pad_left=1, out_rows=64, opf=opf, backend=capi.BACKEND_TCGEN05,
                       segs=[dict(col0=0, ncols=128, res=a[0]["res"], raw=raw)])
        a1 = make_args(a[1]["x"], a[1]["w"], a[1]["bias"], k=7, dil=1, pad_left=3, out_rows=64, opf=opf, backend=capi.BACKEND_TCGEN05,
                       segs=[dict(col0=0, ncols=128, res_op=a[1]["res_op"], res_inv_slope=10.0)])
        arr = (C.POINTER(capi.ConvArgs) * 2)(C.pointer(a0), C.pointer(a1))
        capi.check(lib.qvc_conv1d_sum(arr, 2, stream()), "qvc_conv1d_sum")
