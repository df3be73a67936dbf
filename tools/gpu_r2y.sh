#!/bin/bash
# round 2, call Y: streamed tcgen05 bc 256 training after the ring / read-out / reduction changes: whole training suite + timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_wide.py -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -6 gpurun_out/r2y_pytest.log
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | grep "train phases" | head -1 | cut -c1-900
 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | head -2
 timeout 300 python tools/time_train.py 8192 8192 3 256 2>&1 | head -2
 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 timeout 300 python tools/time_train.py 2048 8192 2 128 2>&1 | head -1) 2>&1 | tee gpurun_out/r2y_time_train.log
