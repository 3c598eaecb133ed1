#!/bin/bash
# chunk-size sweep + launch list + one full ncu capture of the dominant series-convolution kernel
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gpu_tests.log
for c in 2 4 8 16 32 64; do
  python bench.py --steps 5 --warmup 3 --chunk-utts $c --no-extras --no-cpu-baseline > gpurun_out/bench_chunk$c.json 2> gpurun_out/bench_chunk$c.err
done
python bench.py --steps 5 --warmup 3 --chunk-utts 8 --precision bf16 --no-extras --no-cpu-baseline > gpurun_out/bench_bf16_chunk8.json 2>&1
python bench.py --steps 5 --warmup 3 --chunk-utts 64 --precision bf16 --no-extras --no-cpu-baseline > gpurun_out/bench_bf16_chunk64.json 2>&1
python scripts/profile_step.py tf32 8 500 8 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 100 -c 3 -o gpurun_out/r01_conv_tc_full python scripts/profile_step.py tf32 8 500 8 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
