#!/bin/bash
for prec in fp16 tf32; do for ws in 0 6 4 3 2; do QVC_TC2_WS=$ws python scripts/step_time.py $prec 64 500 10; done; done
