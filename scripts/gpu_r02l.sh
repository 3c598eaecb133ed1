#!/bin/bash
# source-level capture of an epilogue-bound pair of MRF-2 layers in fp16 mode: c1 k3 (paired) and its c2
python scripts/profile_step.py fp16 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc2_kernel -s 28 -c 2 -f -o gpurun_out/r02l_mrf2_k3_fp16 python scripts/profile_step.py fp16 64 500 > gpurun_out/ncu_l1.log 2>&1
tail -3 gpurun_out/ncu_l1.log
