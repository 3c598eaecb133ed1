#!/bin/bash
# where does a fused WN layer spend its time?  (QVC_WN_DEBUG experiments; results are garbage, only timings count)
for d in 0 1 2 4 5 7 8 16 24 31; do echo "QVC_WN_DEBUG=$d"; QVC_WN_DEBUG=$d python scripts/wn_bench.py 2>&1 | grep fused; done
