#!/bin/bash
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_infer.py -m gpu -x -q > gpurun_out/r02k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02k_tests.log
tail -5 gpurun_out/r02k_tests.log
for ew in 16 8; do for prec in tf32 fp16; do QVC_EPI_WARPS=$ew python scripts/step_time.py $prec 64 500 10; done; QVC_EPI_WARPS=$ew python scripts/wn_bench.py 2>&1 | grep fused; done
