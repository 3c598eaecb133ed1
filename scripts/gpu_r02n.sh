#!/bin/bash
# re-entry check at HEAD: whole GPU suite, smoke, step times in the three tensor modes, launch list (fp16, tf32), bench line
python -m pytest tests -m gpu -x -q > gpurun_out/r02n_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02n_tests.log
tail -4 gpurun_out/r02n_tests.log
python __graft_entry__.py smoke > gpurun_out/r02n_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02n_smoke.log
for prec in tf32 fp16 bf16; do python scripts/step_time.py $prec 64 500 10; done 2>&1 | tee gpurun_out/r02n_steps.log
python scripts/profile_step.py fp16 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02n_launches_fp16.csv python scripts/profile_step.py fp16 64 500 > gpurun_out/ncu_n1.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02n_launches_tf32.csv python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_n2.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?" >> gpurun_out/r02n_bench.err
tail -3 gpurun_out/r02n_bench.err
