#!/bin/bash
# round 2, call Q: phase profile of the bc 256 (config 3) and bc 128 training kernels
mkdir -p gpurun_out
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | grep "train phases" | head -1 | cut -c1-900
 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | head -2
 LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 2 128 2>&1 | grep "train phases" | head -1 | cut -c1-900
 timeout 300 python tools/time_train.py 2048 8192 2 128 2>&1 | head -1) 2>&1 | tee gpurun_out/r2q_time_train.log
