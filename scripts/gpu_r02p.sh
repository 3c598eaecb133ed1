#!/bin/bash
for m in 0 1 3 7; do for prec in fp16 tf32; do QVC_TC_ROWS=$m timeout 300 python scripts/step_time.py $prec 64 500 10; done; done 2>&1 | grep -v Warn | tee gpurun_out/r02p_steps.log
QVC_TC_ROWS=3 timeout 300 python scripts/wn_bench.py > /dev/null 2>&1 || exit 1
QVC_TC_ROWS=3 ncu --set full --clock-control none --import-source on -k regex:conv_tcr_kernel -s 26 -c 2 -f -o gpurun_out/r02p_rows_fp16 python scripts/wn_bench.py > gpurun_out/ncu_p1.log 2>&1
tail -3 gpurun_out/ncu_p1.log
