#!/bin/bash
# round 2, call K: tcgen05 training step after the batched read-out; phase profiles of both kernels on the same box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "gradients or trajectory or odd_scene or determin" > gpurun_out/r2k_pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_train.log
tail -5 gpurun_out/r2k_pytest_train.log
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep "train phases" | head -1 | cut -c1-700
 LBDRN_TRAIN_H2=1 LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep "train phases" | head -1 | cut -c1-700
 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | head -1
 LBDRN_TRAIN_H2=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1) 2>&1 | tee gpurun_out/r2k_time_train.log
