#!/bin/bash
# round 2, call J: tcgen05 training step (MMA = 3) -- descriptor self-test, training parity suite, timings vs the warp-level kernel
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "m64_images" > gpurun_out/r2j_selftest.log 2>&1; echo "selftest rc=$?" >> gpurun_out/r2j_selftest.log
tail -15 gpurun_out/r2j_selftest.log
timeout 200 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "gradients_match_autograd and k5d2_small" > gpurun_out/r2j_grad.log 2>&1; echo "grad rc=$?" >> gpurun_out/r2j_grad.log
tail -25 gpurun_out/r2j_grad.log
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/r2j_pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_train.log
tail -25 gpurun_out/r2j_pytest_train.log
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep "train phases" | head -1 | cut -c1-700
 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | head -1
 LBDRN_TRAIN_H2=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1) 2>&1 | tee gpurun_out/r2j_time_train.log
