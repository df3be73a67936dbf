#!/bin/bash
# round-1 session-2 GPU pass: bc=256 training parity, wide tensor decode parity + timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gpu_wide.py -x -q 2>&1 | tail -25 > gpurun_out/pytest_wide.log; echo "wide rc=$?"; tail -5 gpurun_out/pytest_wide.log
timeout 900 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -25 > gpurun_out/pytest_train.log; echo "train rc=$?"; tail -5 gpurun_out/pytest_train.log
timeout 900 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -25 > gpurun_out/pytest_decode.log; echo "decode rc=$?"; tail -5 gpurun_out/pytest_decode.log
(timeout 300 python tools/time_decode.py 4096 auto 5 3 256; timeout 300 python tools/time_decode.py 4096 precise 2 3 256; timeout 300 python tools/time_decode.py 8192 auto 5 2 64) 2>&1 | tee gpurun_out/time_wide.log
(timeout 600 python tools/time_train.py 1024 8192 3 256; timeout 300 python tools/time_train.py 2048 8192 2 64) 2>&1 | tee gpurun_out/time_train256.log
