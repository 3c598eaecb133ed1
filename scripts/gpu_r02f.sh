#!/bin/bash
python -m pytest tests/test_gpu_ragged.py tests/test_gpu_convert.py -m gpu -x -q -s > gpurun_out/r02f_ragged.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_ragged.log
tail -12 gpurun_out/r02f_ragged.log
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02f_tests.log
tail -5 gpurun_out/r02f_tests.log
python bench.py --steps 10 --warmup 3 --sweep-utts 0 --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?" >> gpurun_out/r02f_bench.err
