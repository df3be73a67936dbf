#!/bin/bash
# round 2, call R: bc 256 training on the 3xTF32 warp-level path with 32-pixel chunks (was FFMA): parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/r2r_pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest_train.log
tail -5 gpurun_out/r2r_pytest_train.log
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | grep "train phases" | head -1 | cut -c1-900
 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | head -2
 LBDRN_TRAIN_FFMA=1 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | head -1
 timeout 300 python tools/time_train.py 8192 8192 3 256 2>&1 | head -2) 2>&1 | tee gpurun_out/r2r_time_train.log
