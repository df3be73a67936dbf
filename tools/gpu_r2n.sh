#!/bin/bash
# round 2, call N: phase profile with the reduction's sub-phases
mkdir -p gpurun_out
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep "train phases" | head -1 | cut -c1-800
 LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | grep "train phases" | head -1 | cut -c1-800
 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1) 2>&1 | tee gpurun_out/r2n_time_train.log
