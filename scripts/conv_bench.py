"""Per-layer microbenchmark of qvc_conv1d (tcgen05 back end) on one B200.

    python scripts/conv_bench.py [tf32|bf16]

Times representative layers of the path at B = 64 x 10 s under a few tiling overrides (environment
variables read by launch_conv_tc at every launch) and prints us / TFLOP/s / GB/s per configuration.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from gpu_util import conv1d, op_dtype, to_op  # noqa: E402
from quickvc_official_b200 import capi  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
opf = capi.OPF_TF32 if prec == "tf32" else capi.OPF_BF16
E = 4 if prec == "tf32" else 2
DEV = "cuda:0"
B = int(os.environ.get("CB_BATCH", "64"))
ROWS_DIV = int(os.environ.get("CB_ROWS_DIV", "1"))      # 2: the 5 s clip geometry
GRAPH = bool(int(os.environ.get("CB_GRAPH", "0")))      # time 20 launches replayed from a CUDA graph (no host gaps)

# name, rows, cin, cout, k, dil, kind
LAYERS = [
    ("mrf2_k3_c1", 10000, 128, 128, 3, 1, "c1"),
    ("mrf2_k3_c2", 10000, 128, 128, 3, 1, "c2"),
    ("mrf2_k7_c1", 10000, 128, 128, 7, 3, "c1"),
    ("mrf2_k7_c2", 10000, 128, 128, 7, 1, "c2"),
    ("mrf2_k11_c1", 10000, 128, 128, 11, 5, "c1"),
    ("mrf1_k3_c1", 2500, 256, 256, 3, 1, "c1"),
    ("mrf1_k7_c1", 2500, 256, 256, 7, 3, "c1"),
    ("mrf1_k7_c2", 2500, 256, 256, 7, 1, "c2"),
    ("mrf1_k11_c1", 2500, 256, 256, 11, 5, "c1"),
    ("wn_in", 500, 192, 384, 5, 1, "gate"),
    ("wn_rs", 500, 192, 384, 1, 1, "rs"),
    ("ups0", 500, 512, 1280, 4, 1, "c1"),
    ("ups1", 2500, 256, 512, 5, 1, "c1"),
    ("conv_post", 10001, 128, 80, 7, 1, "raw"),
    # experiments: tap shifts that keep the slab descriptor 1024-byte aligned, and a pure GEMM of the same FLOPs
    ("x_k7_d8", 10000, 128, 128, 7, 8, "c1"),
    ("x_k7_d1", 10000, 128, 128, 7, 1, "c1"),
    ("x_k1_c896", 10000, 896, 128, 1, 1, "c1"),
    ("x_k1_c128", 10000, 128, 128, 1, 1, "c1"),
    ("x_k2_d8", 10000, 128, 128, 2, 8, "c1"),
]
only = os.environ.get("CB_ONLY")
CONFIGS = [dict(), dict(QVC_TC_SS="2", QVC_TC_WS="8"), dict(QVC_TC_G="1"), dict(QVC_TC_G="1", QVC_TC_SS="2", QVC_TC_WS="8"),
           dict(QVC_TC_SS="2", QVC_TC_WS="5"), dict(QVC_TC_N="128"), dict(QVC_TC_N="64")]
if os.environ.get("CB_GRIDS"):
    CONFIGS = [dict(QVC_TC_GRID=g) for g in os.environ["CB_GRIDS"].split(",")]
if os.environ.get("CB_DEBUGS"):
    CONFIGS = [dict(QVC_TC_DEBUG=g) for g in os.environ["CB_DEBUGS"].split(",")]
if os.environ.get("CB_CFGS"):     # e.g. "QVC_TC_XPROM=256;QVC_TC_DEBUG=8,QVC_TC_XPROM=256"
    CONFIGS = [dict(kv.split("=") for kv in c.split(",") if kv) for c in os.environ["CB_CFGS"].split(";")]
KEYS = ("QVC_TCR_DEBUG", "QVC_TCR_SS", "QVC_TCR_WS", "QVC_TC_ROWS", "QVC_TC_XPROM", "QVC_TC_DEBUG", "QVC_TC_SS", "QVC_TC_WS", "QVC_TC_G", "QVC_TC_N", "QVC_TC_GRID")


def run_layer(name, rows, cin, cout, k, dil, kind):
    rows = rows // ROWS_DIV
    g = torch.Generator().manual_seed(1)
    x = to_op(torch.randn(B, rows, cin, generator=g), opf).to(DEV)
    w = to_op(torch.randn(cout, k, cin, generator=g) / (cin * k) ** 0.5, opf).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    pad = (k - 1) * dil // 2
    kw = dict(k=k, dil=dil, pad_left=pad, out_rows=rows, opf=opf, backend=capi.BACKEND_TCGEN05)
    nbytes = B * rows * cin * E
    if kind == "c1":
        op = torch.empty(B, rows, cout, device=DEV, dtype=op_dtype(opf))
        segs = [dict(col0=0, ncols=cout, slope=0.1, op=op)]
        nbytes += B * rows * cout * E
        fn = lambda: conv1d(x, w, bias, segs=segs, **kw)
    elif kind == "raw":
        raw = torch.empty(B, rows, cout, device=DEV)
        segs = [dict(col0=0, ncols=cout, raw=raw)]
        nbytes += B * rows * cout * 4
        fn = lambda: conv1d(x, w, bias, segs=segs, **kw)
    elif kind == "c2":
        res = torch.randn(B, rows, cout, device=DEV)
        raw = torch.empty(B, rows, cout, device=DEV)
        op = torch.empty(B, rows, cout, device=DEV, dtype=op_dtype(opf))
        segs = [dict(col0=0, ncols=cout, slope=0.1, res=res, raw=raw, op=op)]
        nbytes += B * rows * cout * (8 + E)
        fn = lambda: conv1d(x, w, bias, segs=segs, **kw)
    elif kind == "gate":
        op = torch.empty(B, rows, cout // 2, device=DEV, dtype=op_dtype(opf))
        segs = [dict(col0=0, ncols=cout // 2, op=op)]
        nbytes += B * rows * cout // 2 * E
        fn = lambda: conv1d(x, w, bias, epilogue=capi.EPI_GATE, segs=segs, **kw)
    else:  # rs: two segments, residual + skip accumulate
        H = cout // 2
        xr = torch.randn(B, rows, H, device=DEV)
        xo = torch.empty(B, rows, H, device=DEV, dtype=op_dtype(opf))
        sk = torch.randn(B, rows, H, device=DEV)
        segs = [dict(col0=0, ncols=H, res=xr, raw=xr, op=xo), dict(col0=H, ncols=H, accin=sk, raw=sk)]
        nbytes += B * rows * H * (8 + E + 8)
        fn = lambda: conv1d(x, w, bias, segs=segs, **kw)
    flops = 2.0 * B * rows * cin * cout * k
    for cfg in CONFIGS:
        for kk in KEYS:
            os.environ.pop(kk, None)
        os.environ.update(cfg)
        try:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            n = 20 if GRAPH else 10
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if GRAPH:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    fn()
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    for _ in range(n):
                        fn()
                graph.replay()
                torch.cuda.synchronize()
                s.record()
                graph.replay()
                e.record()
            else:
                s.record()
                for _ in range(n):
                    fn()
                e.record()
            torch.cuda.synchronize()
            us = s.elapsed_time(e) * 1e3 / n
            print(f"{name:12s} {prec} {str(cfg):60s} {us:9.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  {nbytes / us / 1e3:7.1f} GB/s", flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f"{name:12s} {prec} {str(cfg):60s} FAILED {str(ex)[:100]}", flush=True)
    for kk in KEYS:
        os.environ.pop(kk, None)


for L in LAYERS:
    if only and only not in L[0]:
        continue
    run_layer(*L)
