#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r03k_tests.log
tail -4 gpurun_out/r03k_tests.log
python __graft_entry__.py smoke > gpurun_out/r03k_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r03k_smoke.log; tail -5 gpurun_out/r03k_smoke.log
/usr/bin/time -v python bench.py > gpurun_out/r03k_bench.json 2> gpurun_out/r03k_bench.err; echo "bench rc=$?" >> gpurun_out/r03k_bench.err
grep -E "Elapsed|bench rc" gpurun_out/r03k_bench.err
