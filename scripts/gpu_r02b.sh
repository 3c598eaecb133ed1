#!/bin/bash
# round 2, call B: sum-of-convolutions MRF -- kernel tests first, then the whole suite, then bench lines
set -x
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "conv_sum" > gpurun_out/r02b_sum.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_sum.log
tail -15 gpurun_out/r02b_sum.log
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02b_tests.log
tail -8 gpurun_out/r02b_tests.log
python bench.py --steps 10 --warmup 3 --sweep-utts 0 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?" >> gpurun_out/r02b_bench.err
tail -3 gpurun_out/r02b_bench.err
