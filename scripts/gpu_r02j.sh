#!/bin/bash
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "residual_from_operand" > gpurun_out/r02j_k.log 2>&1; echo "rc=$?" >> gpurun_out/r02j_k.log; tail -6 gpurun_out/r02j_k.log
python -m pytest tests -m gpu -x -q > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02j_tests.log
tail -5 gpurun_out/r02j_tests.log
for s in 1 0; do QVC_RES_STAGE=$s python scripts/step_time.py fp16 64 500 10; QVC_RES_STAGE=$s python scripts/step_time.py bf16 64 500 10; done
python scripts/profile_step.py fp16 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02j_launches_fp16.csv python scripts/profile_step.py fp16 64 500 > gpurun_out/ncu_j1.log 2>&1
