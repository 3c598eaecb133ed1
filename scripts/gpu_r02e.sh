#!/bin/bash
for prec in tf32 fp16; do for c in 0 37 32 18 16 8; do QVC_WN_CHUNK=$c python scripts/step_time.py $prec 64 500 10; done; done
