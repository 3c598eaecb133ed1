#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r03m_tests.log
tail -3 gpurun_out/r03m_tests.log
timeout 200 python scripts/latency.py 2>&1 | grep p50 | grep cached
for prec in fp16 tf32 bf16; do timeout 300 python scripts/step_time.py $prec 64 500 20; done 2>&1 | grep -v Warn
