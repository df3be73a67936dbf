#!/bin/bash
# round 2, call M: line-coalesced cross-CTA reduction (all training kernels) + tcgen05 step: full training suite, timings, phases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/r2m_pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_train.log
tail -5 gpurun_out/r2m_pytest_train.log
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep "train phases" | head -1 | cut -c1-700
 LBDRN_TRAIN_H2=1 LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep "train phases" | head -1 | cut -c1-700
 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | head -1
 LBDRN_TRAIN_H2=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | head -1
 timeout 300 python tools/time_train.py 2048 8192 2 128 2>&1 | head -1) 2>&1 | tee gpurun_out/r2m_time_train.log
