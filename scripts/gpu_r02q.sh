#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "rows_kernel or fused_wn" > gpurun_out/r02q_k.log 2>&1; echo "rc=$?" >> gpurun_out/r02q_k.log; tail -5 gpurun_out/r02q_k.log
grep -q "rc=0" gpurun_out/r02q_k.log || exit 1
(echo "== rows 3"; timeout 300 python scripts/wn_bench.py) 2>&1 | grep -v Warn | tee gpurun_out/r02q_wn.log
for m in 1 3 7; do for prec in fp16 tf32; do QVC_TC_ROWS=$m timeout 300 python scripts/step_time.py $prec 64 500 20; done; done 2>&1 | grep -v Warn | tee gpurun_out/r02q_steps.log
