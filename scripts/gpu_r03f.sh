#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_infer.py -m gpu -x -q -k "fused_post_net or golden or tail" > gpurun_out/r03f_k.log 2>&1; echo "rc=$?" >> gpurun_out/r03f_k.log; tail -4 gpurun_out/r03f_k.log
grep -q "rc=0" gpurun_out/r03f_k.log || exit 1
python scripts/profile_step.py fp16 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv -k regex:post_tail --log-file gpurun_out/r03f_pt.csv python scripts/profile_step.py fp16 64 500 > gpurun_out/ncu_f1.log 2>&1
grep post_tail gpurun_out/r03f_pt.csv | tail -1
