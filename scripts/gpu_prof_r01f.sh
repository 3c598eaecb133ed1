#!/bin/bash
# current kernels: launch list (tf32 + bf16) and full ncu captures of representative layers
mkdir -p gpurun_out
python scripts/profile_step.py tf32 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01f_tf32.csv --profile-from-start off python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_f1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01f_bf16.csv --profile-from-start off python scripts/profile_step.py bf16 64 500 > gpurun_out/ncu_f2.log 2>&1
for spec in "gate:1:2" "mrf1k11:88:2" "mrf2k3:95:2" "mrf2k7:101:2"; do
  name=${spec%%:*}; rest=${spec#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --profile-from-start off -s $skip -c $cnt \
      -o gpurun_out/r01f_$name -f python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_f_$name.log 2>&1
  tail -1 gpurun_out/ncu_f_$name.log
done
ncu --set full --clock-control none --import-source on -k regex:tail_kernel --profile-from-start off -c 1 -o gpurun_out/r01f_tail -f python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_f_tail.log 2>&1
