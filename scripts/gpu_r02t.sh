#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02t_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02t_tests.log
tail -5 gpurun_out/r02t_tests.log
for cfg in "QVC_WN_DEFER=0 QVC_TC_ROWS=1" "QVC_WN_DEFER=1 QVC_TC_ROWS=1" "QVC_WN_DEFER=1 QVC_TC_ROWS=3"; do for prec in fp16 tf32; do env $cfg timeout 300 python scripts/step_time.py $prec 64 500 20; done; done 2>&1 | grep -v Warn | tee gpurun_out/r02t_steps.log
python scripts/profile_step.py fp16 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02t_launches_fp16.csv python scripts/profile_step.py fp16 64 500 > gpurun_out/ncu_t1.log 2>&1
