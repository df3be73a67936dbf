#!/bin/bash
mkdir -p gpurun_out
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep "train phases" | head -1 | cut -c1-700
 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | head -1) | tee gpurun_out/time_train.log
