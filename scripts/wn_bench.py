"""Fused WN layer (qvc_wn_layer) against the two separate convolutions, B = 64 x 10 s."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from gpu_util import make_args, op_dtype, stream, to_op  # noqa: E402
from quickvc_official_b200 import capi  # noqa: E402

os.environ.setdefault("QVC_WN_FUSED", "1")          # qvc_wn_layer is off by default (the row kernel serves the WN stacks)
DEV = "cuda:0"
lib = capi.load()
for prec, opf in (("tf32", capi.OPF_TF32), ("bf16", capi.OPF_BF16), ("fp16", capi.OPF_F16)):
    B, rows, H, k = int(os.environ.get("CB_BATCH", "64")), 500, 192, 5
    g = torch.Generator().manual_seed(1)
    x = to_op(torch.randn(B, rows, H, generator=g), opf).to(DEV)
    w_in = to_op(torch.randn(2 * H, k, H, generator=g) / (H * k) ** 0.5, opf).to(DEV)
    gbias = torch.randn(1, 2 * H, generator=g).to(DEV)
    w_rs = to_op(torch.randn(2 * H, 1, H, generator=g) / H ** 0.5, opf).to(DEV)
    b_rs = torch.randn(2 * H, generator=g).to(DEV)
    xr = torch.randn(B, rows, H, generator=g).to(DEV)
    sk = torch.randn(B, rows, H, generator=g).to(DEV)
    acts = torch.zeros(B, rows, H, device=DEV, dtype=op_dtype(opf))
    xo = torch.zeros(B, rows, H, device=DEV, dtype=op_dtype(opf))
    be = capi.BACKEND_TCGEN05
    a = make_args(x, w_in, gbias, k=k, dil=1, pad_left=2, out_rows=rows, opf=opf, backend=be, epilogue=capi.EPI_GATE,
                  segs=[dict(col0=0, ncols=H, op=acts)])
    r = make_args(acts, w_rs, b_rs, k=1, dil=1, pad_left=0, out_rows=rows, opf=opf, backend=be,
                  segs=[dict(col0=0, ncols=H, res=xr, raw=xr, op=xo), dict(col0=H, ncols=H, accin=sk, raw=sk)])

    def fused():
        capi.check(lib.qvc_wn_layer(C.byref(a), C.byref(r), stream()), "qvc_wn_layer")

    def split():
        capi.check(lib.qvc_conv1d(C.byref(a), stream()), "qvc_conv1d")
        capi.check(lib.qvc_conv1d(C.byref(r), stream()), "qvc_conv1d")

    def gate_only():
        capi.check(lib.qvc_conv1d(C.byref(a), stream()), "qvc_conv1d")

    def rs_only():
        capi.check(lib.qvc_conv1d(C.byref(r), stream()), "qvc_conv1d")

    for name, fn in (("fused", fused), ("split", split), ("gate", gate_only), ("rs", rs_only)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        print(f"{prec} {name:6s} {s.elapsed_time(e) * 1e3 / n:8.1f} us per layer", flush=True)
