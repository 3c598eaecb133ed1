#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "rows_kernel or fused_wn" > gpurun_out/r02r_k.log 2>&1; echo "rc=$?" >> gpurun_out/r02r_k.log; tail -5 gpurun_out/r02r_k.log
grep -q "rc=0" gpurun_out/r02r_k.log || exit 1
(echo "== rows 3"; timeout 300 python scripts/wn_bench.py) 2>&1 | grep -v Warn | tee gpurun_out/r02r_wn.log
