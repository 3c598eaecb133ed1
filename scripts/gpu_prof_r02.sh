#!/bin/bash
# round-2 profile set: launch lists in the three tensor modes, DRAM traffic per launch (tf32, fp16), full captures of the
# dominant kernels (source-level), then the default bench and the reference arm.  Every ncu pass follows a plain run of the
# same command that exited 0.
mkdir -p gpurun_out
for prec in tf32 fp16 bf16; do python scripts/profile_step.py $prec 64 500 >> gpurun_out/prof_plain.log 2>&1 || exit 1; done
NCU="ncu --clock-control none --profile-from-start off"
for prec in tf32 fp16 bf16; do
  $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02z_launches_$prec.csv python scripts/profile_step.py $prec 64 500 > gpurun_out/ncu_z_$prec.log 2>&1
done
for prec in tf32 fp16; do
  $NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/r02z_traffic_$prec.csv python scripts/profile_step.py $prec 64 500 > gpurun_out/ncu_zt_$prec.log 2>&1
done
# full captures (tf32 step): the WN gate layer and a residual layer on the row kernel, the two MRF-2 k = 3 layers (c1, c2: memory-bound) on it, an MRF-1 k = 11
# layer on the channel-major pair kernel, the fused post-net + tail kernel
$NCU --set full --import-source on -k regex:conv_tcr_kernel -s 1 -c 2 -f -o gpurun_out/r02z_tcr_wn_tf32 python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_z1.log 2>&1
$NCU --set full --import-source on -k regex:conv_tcr_kernel -s 73 -c 2 -f -o gpurun_out/r02z_tcr_mrf2_k3_c1c2_tf32 python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_z2.log 2>&1
$NCU --set full --import-source on -k regex:conv_tc2_kernel -s 12 -c 1 -f -o gpurun_out/r02z_tc2_mrf1_k11_tf32 python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_z3.log 2>&1
$NCU --set full --import-source on -k regex:post_tail_kernel -c 1 -f -o gpurun_out/r02z_post_tail_tf32 python scripts/profile_step.py tf32 64 500 > gpurun_out/ncu_z4.log 2>&1
python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?" >> gpurun_out/r02z_bench.err
python bench.py --precision fp16 --sweep-utts 0 --no-cpu-baseline > gpurun_out/r02z_bench_fp16.json 2> gpurun_out/r02z_bench_fp16.err; echo "bench rc=$?" >> gpurun_out/r02z_bench_fp16.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_bench_ref.json 2> gpurun_out/r02z_bench_ref.err
tail -2 gpurun_out/r02z_bench.err gpurun_out/r02z_bench_fp16.err
