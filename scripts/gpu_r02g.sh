#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02g_tests.log
tail -5 gpurun_out/r02g_tests.log
python bench.py --steps 10 --warmup 3 --sweep-utts 0 --no-cpu-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?" >> gpurun_out/r02g_bench.err
python scripts/profile_step.py fp16 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02g_launches_fp16.csv python scripts/profile_step.py fp16 64 500 > gpurun_out/ncu_g1.log 2>&1
