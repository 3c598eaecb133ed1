#!/bin/bash
# v2 kernel: tests, bench, and full ncu captures of representative layers (conv_tc launch ordinals of one step)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
python bench.py --steps 5 --warmup 3 --precision bf16 --no-extras --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
python scripts/profile_step.py tf32 32 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
for spec in "gate:1:2" "mrf1k7:82:2" "mrf2k3:95:2" "mrf2k7:101:2"; do
  name=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --profile-from-start off -s $skip -c $cnt \
      -o gpurun_out/r01c_$name -f python scripts/profile_step.py tf32 32 500 > gpurun_out/ncu_$name.log 2>&1
  tail -2 gpurun_out/ncu_$name.log
done
