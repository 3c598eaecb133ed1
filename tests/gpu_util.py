"""Helpers for the -m gpu tests: call the C ABI on torch CUDA tensors."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from quickvc_official_b200 import capi, fold


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def tref(t: Optional[torch.Tensor], bstride=None) -> capi.Tensor:
    """[B][rows][ch] contiguous tensor -> qvc_tensor."""
    if t is None:
        return capi.Tensor(None, 0, 0, 0)
    assert t.is_contiguous() and t.dim() == 3
    return capi.Tensor(t.data_ptr(), t.shape[1] * t.shape[2] if bstride is None else bstride, t.shape[2], 0)


def op_dtype(opf: int):
    return {capi.OPF_BF16: torch.bfloat16, capi.OPF_F16: torch.float16}.get(opf, torch.float32)


def to_op(x: torch.Tensor, opf: int) -> torch.Tensor:
    return fold.to_operand(x, opf)


def make_args(x, w, bias, *, k, dil, pad_left, out_rows, opf, backend, epilogue=capi.EPI_LINEAR, segs=(),
              noise=None, aux0=None, aux1=None, bias_bstride=0, taps=None) -> capi.ConvArgs:
    """x [B][rows][cin] operand tensor, w [cout][k][cin] operand tensor; segs: list of dicts with
    col0, ncols, alpha, beta, slope, res, accin, raw, op tensors.  The caller keeps the tensors alive."""
    a = capi.ConvArgs()
    a.x = tref(x)
    a.batch, a.x_rows, a.out_rows, a.cin = x.shape[0], x.shape[1], out_rows, x.shape[2]
    a.w = w.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    a.bias_bstride = bias_bstride
    a.cout, a.k, a.dil, a.pad_left = w.shape[0], k, dil, pad_left
    a.epilogue, a.nseg = epilogue, len(segs)
    for i, s in enumerate(segs):
        g = a.seg[i]
        g.col0, g.ncols = s["col0"], s["ncols"]
        g.alpha, g.beta, g.slope = s.get("alpha", 1.0), s.get("beta", 1.0), s.get("slope", 1.0)
        g.res, g.accin, g.raw, g.op = (tref(s.get(n)) for n in ("res", "accin", "raw", "op"))
        g.res_op, g.res_inv_slope = tref(s.get("res_op")), s.get("res_inv_slope", 1.0)
    a.noise, a.aux0, a.aux1 = tref(noise), tref(aux0), tref(aux1)
    a.opformat, a.backend = opf, backend
    if taps is not None:            # structured-zero hint: (tap_split, lo[2][2], hi[2][2])
        a.tap_split = taps[0]
        for p in range(2):
            for q in range(2):
                a.tap_lo[p][q], a.tap_hi[p][q] = taps[1][p][q], taps[2][p][q]
    return a


def conv1d(x, w, bias, **kw):
    a = make_args(x, w, bias, **kw)
    capi.check(capi.load().qvc_conv1d(C.byref(a), stream()), "qvc_conv1d")


def ref_conv(x, w, k, dil, pad_left, out_rows):
    """fp64 reference of the series convolution on (already rounded) operand tensors."""
    import torch.nn.functional as F
    rows = x.shape[1]
    right = (out_rows - 1) + (k - 1) * dil - pad_left - (rows - 1)
    xt = F.pad(x.double().transpose(1, 2), (pad_left, max(right, 0)))
    y = F.conv1d(xt, w.double().permute(0, 2, 1), dilation=dil)[:, :, :out_rows]
    return y.transpose(1, 2)
