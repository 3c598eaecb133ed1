#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "rows_kernel" > gpurun_out/r02z_k.log 2>&1; echo "rc=$?" >> gpurun_out/r02z_k.log; tail -5 gpurun_out/r02z_k.log
grep -q "rc=0" gpurun_out/r02z_k.log || exit 1
for p in bf16 tf32; do CB_CFGS="QVC_TC_ROWS=0;QVC_TC_ROWS=7" timeout 280 python scripts/conv_bench.py $p 2>&1 | grep -v Warn | grep -E "mrf|wn_|conv_post|ups"; done | tee gpurun_out/r02z_conv.log
for prec in fp16 tf32; do timeout 300 python scripts/step_time.py $prec 64 500 20; done 2>&1 | grep -v Warn | tee gpurun_out/r02z_steps.log
