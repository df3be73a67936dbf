#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -x -q 2>&1 | tail -5 > gpurun_out/pytest_decode.log; echo "decode rc=$?"; tail -3 gpurun_out/pytest_decode.log
(timeout 300 python tools/time_decode.py 8192 auto 10 2 64; timeout 300 python tools/time_decode.py 8192 tensor 10 2 64; LBDRN_TC_NWG1=1 timeout 300 python tools/time_decode.py 8192 auto 10 2 64) 2>&1 | tee gpurun_out/time_tc64.log
