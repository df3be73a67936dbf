#!/bin/bash
# round 2, call X: streamed tcgen05 training variant for bc 256 (MMA = 4): parity + timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "d3_bc256" > gpurun_out/r2x_pytest_bc256.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest_bc256.log
tail -30 gpurun_out/r2x_pytest_bc256.log
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | grep "train phases" | head -1 | cut -c1-900
 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | head -2
 LBDRN_TRAIN_TF32=1 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | head -1) 2>&1 | tee gpurun_out/r2x_time_train.log
