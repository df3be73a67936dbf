#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -8 > gpurun_out/pytest_train.log; echo "train rc=$?"; tail -3 gpurun_out/pytest_train.log
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | tail -8; timeout 300 python tools/time_train.py 2048 8192 2 64 | head -1) 2>&1 | tee gpurun_out/time_train.log
