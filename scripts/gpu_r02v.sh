#!/bin/bash
# what bounds a memory-bound 128 -> 128 k = 3 layer (MRF-2 c1) on the row kernel and on the channel-major kernel?
export CB_ONLY=mrf2_k3_c1
CB_CFGS="QVC_TC_ROWS=7" ncu --set full --clock-control none --import-source on -k regex:conv_tcr_kernel -s 5 -c 1 -f -o gpurun_out/r02v_rows_k3c1_bf16 python scripts/conv_bench.py bf16 > gpurun_out/ncu_v1.log 2>&1
CB_CFGS="QVC_TC_ROWS=0" ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 5 -c 1 -f -o gpurun_out/r02v_x_k3c1_bf16 python scripts/conv_bench.py bf16 > gpurun_out/ncu_v2.log 2>&1
tail -2 gpurun_out/ncu_v1.log gpurun_out/ncu_v2.log
