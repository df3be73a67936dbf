#!/bin/bash
# round 2, call 3g: where the generic gather's cycles go (sub-phase markers), bc 256 / bc 128 / bc 32
mkdir -p gpurun_out
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | grep "train phases" | head -1 | cut -c1-1000
 LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 2 128 2>&1 | grep "train phases" | head -1 | cut -c1-1000
 LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 2 32 2>&1 | grep "train phases" | head -1 | cut -c1-1000
 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | head -1) 2>&1 | tee gpurun_out/r3g_time_train.log
