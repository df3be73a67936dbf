#!/bin/bash
# decode parity tests (all decode kernels) + headline timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_fullsize.py tests/test_gpu_wide.py -x -q > gpurun_out/pytest_decode.log 2>&1; echo "decode rc=$?"
grep -v "^frame\|^#" gpurun_out/pytest_decode.log | tail -3
(timeout 200 python tools/time_decode.py 8192 auto 5 2 64; timeout 200 python tools/time_decode.py 8192 auto 5 2 64) 2>&1 | tee gpurun_out/time_tc64.log
