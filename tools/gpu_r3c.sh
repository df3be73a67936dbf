#!/bin/bash
# round 2: low-order-weight decode on the wide kernel (sibling launch): wide + decode suites, config-3 decode timing unchanged?
mkdir -p gpurun_out
LBDRN_DEBUG=1 timeout 300 python -m pytest tests/test_gpu_wide.py -m gpu -x -q -k "full_precision" > gpurun_out/r3c_pytest_wlo.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3c_pytest_wlo.log
tail -25 gpurun_out/r3c_pytest_wlo.log
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_decode.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r3c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3c_pytest.log
tail -5 gpurun_out/r3c_pytest.log
timeout 300 python tools/time_decode.py 2>&1 | tail -12 | tee gpurun_out/r3c_time_decode.log
