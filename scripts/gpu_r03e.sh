#!/bin/bash
python scripts/profile_step.py fp16 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r03e_launches_fp16.csv python scripts/profile_step.py fp16 64 500 > gpurun_out/ncu_e1.log 2>&1
$NCU --set full --import-source on -k regex:post_tail_kernel -c 1 -f -o gpurun_out/r03e_post_tail_fp16 python scripts/profile_step.py fp16 64 500 > gpurun_out/ncu_e2.log 2>&1
$NCU --set full --import-source on -k regex:post_tail_kernel -c 1 -f -o gpurun_out/r03e_post_tail_tf32 python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_e3.log 2>&1
