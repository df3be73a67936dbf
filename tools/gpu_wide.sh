#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wide.py -x -q 2>&1 | tail -25 > gpurun_out/pytest_wide.log; echo "wide rc=$?"; tail -5 gpurun_out/pytest_wide.log
(LBDRN_TCW_PROF=1 timeout 300 python tools/time_decode.py 4096 auto 1 3 256 2>&1 | tail -6
 timeout 300 python tools/time_decode.py 4096 auto 5 3 256; timeout 300 python tools/time_decode.py 4096 auto 5 2 128;  timeout 300 python tools/time_decode.py 8192 auto 10 2 64) 2>&1 | tee gpurun_out/time_wide.log
