#!/bin/bash
# e2e robustness check + per-launch lists of the sum-of-convolutions MRF (tf32, fp16)
python bench.py --steps 10 --warmup 3 --sweep-utts 0 --no-cpu-baseline --no-extras > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
QVC_BENCH_NO_SAMPLER=1 python bench.py --steps 10 --warmup 3 --sweep-utts 0 --no-cpu-baseline --no-extras > gpurun_out/r02d_bench_nosampler.json 2>> gpurun_out/r02d_bench.err
python scripts/profile_step.py tf32 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
python scripts/profile_step.py fp16 64 500 >> gpurun_out/prof_plain.log 2>&1 || exit 1
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02d_launches_tf32.csv python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_d1.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02d_launches_fp16.csv python scripts/profile_step.py fp16 64 500 > gpurun_out/ncu_d2.log 2>&1
