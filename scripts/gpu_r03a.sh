#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r03a_tests.log
tail -5 gpurun_out/r03a_tests.log
for m in 3 11; do for prec in fp16 tf32 bf16; do QVC_TC_ROWS=$m timeout 300 python scripts/step_time.py $prec 64 500 20; done; done 2>&1 | grep -v Warn | tee gpurun_out/r03a_steps.log
python scripts/profile_step.py fp16 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r03a_launches_fp16.csv python scripts/profile_step.py fp16 64 500 > gpurun_out/ncu_a1.log 2>&1
ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r03a_launches_tf32.csv python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_a2.log 2>&1
