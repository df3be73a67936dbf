#!/bin/bash
# round 2, call 3i: phase profile of the bc 64 tcgen05 step (paper configuration) and of the window-3 / window-7 variants
mkdir -p gpurun_out
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep "train phases" | head -1 | cut -c1-1000
 LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 1 64 2>&1 | grep "train phases" | head -1 | cut -c1-1000
 LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 3 64 2>&1 | grep "train phases" | head -1 | cut -c1-1000
 timeout 300 python tools/time_train.py 2048 8192 3 64 2>&1 | head -1) 2>&1 | tee gpurun_out/r3i_time_train.log
