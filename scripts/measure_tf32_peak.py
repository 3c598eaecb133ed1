"""Measured TF32 tensor peak on this pool's B200, the same way the driver measured MEASURED_PEAKS.json's bf16 figures:
torch.matmul (cuBLAS) 8192^3 with TF32 operands, best of 10 (burst) and back to back for 4 s (sustained).  SURVEY.md
section 8d leaves this number to the builder; bench.py uses it as the roofline denominator of the fp32 (TF32) mode."""
import json
import time

import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
c = torch.empty(n, n, device="cuda")
flops = 2.0 * n ** 3
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); torch.matmul(a, b, out=c); e.record(); torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e))
t0 = time.time()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
k = 0
while time.time() - t0 < 4.0:
    for _ in range(20):
        torch.matmul(a, b, out=c)
    k += 20
    torch.cuda.synchronize()
e.record(); torch.cuda.synchronize()
out = {"tf32_tflops": flops / (best * 1e-3) / 1e12, "tf32_tflops_sustained": flops * k / (s.elapsed_time(e) * 1e-3) / 1e12,
       "how": "torch.matmul fp32 8192^3 with torch.backends.cuda.matmul.allow_tf32 = True (cuBLAS TF32): best of 10 (burst) and "
              "back to back for 4 s (sustained), CUDA events", "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
print(json.dumps(out))
