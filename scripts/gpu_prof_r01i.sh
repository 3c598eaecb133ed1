#!/bin/bash
# round-1 final code (ABI v4): launch lists in both modes, DRAM traffic of one step, full captures of one launch of each
# tensor-core kernel (fused WN layer, CTA-pair MRF-1 k11, cta_group::1 MRF-2 k11) and of the tail
mkdir -p gpurun_out
python scripts/profile_step.py tf32 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
python scripts/profile_step.py bf16 64 500 >> gpurun_out/prof_plain.log 2>&1 || exit 1
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r01i_launches_tf32.csv python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_i1.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r01i_launches_bf16.csv python scripts/profile_step.py bf16 64 500 > gpurun_out/ncu_i2.log 2>&1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/r01i_traffic_tf32.csv python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_i3.log 2>&1
$NCU --set full --import-source on -k regex:conv_wn_kernel -s 3 -c 1 -o gpurun_out/r01i_wn -f python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_i4.log 2>&1
$NCU --set full --import-source on -k regex:conv_tc2_kernel -s 10 -c 1 -o gpurun_out/r01i_mrf1 -f python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_i5.log 2>&1
$NCU --set full --import-source on -k regex:conv_tc_kernel -s 30 -c 2 -o gpurun_out/r01i_mrf2 -f python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_i6.log 2>&1
$NCU --set full --import-source on -k regex:tail_kernel -c 1 -o gpurun_out/r01i_tail -f python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_i7.log 2>&1
tail -2 gpurun_out/ncu_i7.log
