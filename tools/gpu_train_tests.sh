#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/pytest_train_full.log 2>&1; echo "train rc=$?"
grep -v "^frame\|^#" gpurun_out/pytest_train_full.log | tail -3
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep -v "^Traceback\|File\|print\|Broken" | head -3 | cut -c1-600
 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | head -1) | tee gpurun_out/time_train.log
