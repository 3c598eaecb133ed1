#!/bin/bash
# final round-1 kernels: full ncu captures of the CTA-pair kernel (MRF-1 k11, WN gate), the LSTM and the launch lists
mkdir -p gpurun_out
python scripts/profile_step.py tf32 64 500 > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01h_tf32.csv --profile-from-start off python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_h1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01h_bf16.csv --profile-from-start off python scripts/profile_step.py bf16 64 500 > gpurun_out/ncu_h2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel --profile-from-start off -s 0 -c 1 -o gpurun_out/r01h_gate2 -f python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_h_gate2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel --profile-from-start off -s 61 -c 2 -o gpurun_out/r01h_mrf1k11_2 -f python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_h_mrf1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_recurrent --profile-from-start off -c 1 -o gpurun_out/r01h_lstm -f python scripts/profile_step.py tf32 1 250 > gpurun_out/ncu_h_lstm.log 2>&1
tail -1 gpurun_out/ncu_h_lstm.log
