#!/bin/bash
# frames-on-rows pair kernel (conv_tcr.cu): kernel tests, WN layer timing old vs new, whole step
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "rows_kernel or fused_wn" > gpurun_out/r02o_k.log 2>&1; echo "rc=$?" >> gpurun_out/r02o_k.log; tail -15 gpurun_out/r02o_k.log
grep -q "rc=0" gpurun_out/r02o_k.log || exit 1
(echo "== rows (default)"; timeout 300 python scripts/wn_bench.py; echo "== QVC_TC_ROWS=0"; QVC_TC_ROWS=0 timeout 300 python scripts/wn_bench.py) 2>&1 | grep -v Warn | tee gpurun_out/r02o_wn.log
for prec in tf32 fp16 bf16; do timeout 300 python scripts/step_time.py $prec 64 500 10; done 2>&1 | tee gpurun_out/r02o_steps.log
