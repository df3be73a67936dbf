#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_decode.log; echo "decode rc=$?"; tail -6 gpurun_out/pytest_decode.log
(timeout 300 python tools/time_decode.py 8192 auto 10 2 64
 timeout 300 python tools/time_coords.py 4096 1 auto; timeout 300 python tools/time_coords.py 4096 0 auto; timeout 300 python tools/time_coords.py 4096 1 precise) 2>&1 | tee gpurun_out/time_tc64.log
