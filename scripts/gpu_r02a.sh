#!/bin/bash
# round 2, call A: whole GPU suite on the new tests + bench line with the new legs
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02a_tests.log
python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02a_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?" >> gpurun_out/r02a_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err
tail -3 gpurun_out/r02a_tests.log
