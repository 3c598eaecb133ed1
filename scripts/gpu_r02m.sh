#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r02m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02m_tests.log
tail -4 gpurun_out/r02m_tests.log
for prec in tf32 fp16 bf16; do python scripts/step_time.py $prec 64 500 10; done
for d in 0 32; do echo "QVC_WN_DEBUG=$d"; QVC_WN_DEBUG=$d python scripts/wn_bench.py 2>&1 | grep fused; done
QVC_WN_DEBUG=32 python scripts/step_time.py tf32 64 500 10
QVC_WN_DEBUG=32 python scripts/step_time.py fp16 64 500 10
