#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_post_net or tail_matches" > gpurun_out/r03d_k.log 2>&1; echo "rc=$?" >> gpurun_out/r03d_k.log; tail -30 gpurun_out/r03d_k.log
grep -q "rc=0" gpurun_out/r03d_k.log || exit 1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r03d_tests.log
tail -5 gpurun_out/r03d_tests.log
for pt in 0 1; do for prec in fp16 tf32 bf16; do QVC_POST_TAIL=$pt timeout 300 python scripts/step_time.py $prec 64 500 20; done; done 2>&1 | grep -v Warn | tee gpurun_out/r03d_steps.log
