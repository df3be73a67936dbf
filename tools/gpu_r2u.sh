#!/bin/bash
# round 2, call U: reference sampler through lbdrn_host_randperm32 + side-stream uploads: training suite, encode wall, eval at 2048^2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/r2u_pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest_train.log
tail -4 gpurun_out/r2u_pytest_train.log
(timeout 600 python tools/enc_gap.py 2>&1 | tail -6
 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | head -5) 2>&1 | tee gpurun_out/r2u_enc.log
