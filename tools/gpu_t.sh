#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_train.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_t.log; echo "rc=$?"; tail -12 gpurun_out/pytest_t.log
