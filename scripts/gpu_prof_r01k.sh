#!/bin/bash
# round-1 final code (frame-paired MRF-2, st.async LSTM): launch lists in both modes + DRAM traffic + one capture of a
# frame-paired MRF-2 launch and of the LSTM, then the default bench
mkdir -p gpurun_out
python scripts/profile_step.py tf32 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
python scripts/profile_step.py bf16 64 500 >> gpurun_out/prof_plain.log 2>&1 || exit 1
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r01k_launches_tf32.csv python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_k1.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r01k_launches_bf16.csv python scripts/profile_step.py bf16 64 500 > gpurun_out/ncu_k2.log 2>&1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/r01k_traffic_tf32.csv python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_k3.log 2>&1
$NCU --set full --import-source on -k regex:conv_tc2_kernel -s 40 -c 1 -o gpurun_out/r01k_mrf2pair -f python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_k4.log 2>&1
$NCU --set full --import-source on -k regex:lstm_recurrent -s 1 -c 1 -o gpurun_out/r01k_lstm -f python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_k5.log 2>&1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 400 gpurun_out/bench_final.json
