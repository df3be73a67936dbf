#!/bin/bash
# round 2, call 3f: batched neighbourhood gather on the generic path: training + wide suites, timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_wide.py -m gpu -x -q > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3f_pytest.log
tail -6 gpurun_out/r3f_pytest.log
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | grep "train phases" | head -1 | cut -c1-900
 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | head -2
 timeout 300 python tools/time_train.py 8192 8192 3 256 2>&1 | head -2
 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 timeout 300 python tools/time_train.py 2048 8192 2 128 2>&1 | head -1) 2>&1 | tee gpurun_out/r3f_time_train.log
(timeout 300 python tools/time_train.py 2048 8192 2 32 2>&1 | head -1
 timeout 300 python tools/time_train.py 2048 8192 1 64 2>&1 | head -1) 2>&1 | tee -a gpurun_out/r3f_time_train.log
