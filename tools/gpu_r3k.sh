#!/bin/bash
# round 2, call 3k: wide kernel with the output operand in one ring slot: wide suite + role profile + timings (bc 256 / bc 128, eval)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py -m gpu -x -q > gpurun_out/r3k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3k_pytest.log
tail -3 gpurun_out/r3k_pytest.log
(LBDRN_TCW_PROF=1 timeout 300 python tools/time_decode.py 4096 auto 1 3 256 2>&1 | grep "lbdrn" | head -5
 timeout 300 python tools/time_decode.py 8192 auto 5 3 256 2>&1 | tail -1
 timeout 300 python tools/time_decode.py 8192 auto 5 2 128 2>&1 | tail -1
 timeout 300 python tools/time_train.py 8192 8192 3 256 2>&1 | grep "eval pass" | tail -1) 2>&1 | tee gpurun_out/r3k_tcw.log
