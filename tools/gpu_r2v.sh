#!/bin/bash
# round 2, call V: full-scene evaluation at bc 128 / 256 on the wide tcgen05 kernel (low-order weight operands): parity + timing
mkdir -p gpurun_out
LBDRN_DEBUG=1 timeout 300 python -m pytest tests/test_gpu_wide.py -m gpu -x -q -k "evaluation" > gpurun_out/r2v_pytest_eval.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest_eval.log
tail -25 gpurun_out/r2v_pytest_eval.log
timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_train.py -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -5 gpurun_out/r2v_pytest.log
(timeout 300 python tools/time_train.py 8192 8192 3 256 2>&1 | head -4
 LBDRN_EVAL_FP32=1 timeout 300 python tools/time_train.py 2048 8192 3 256 2>&1 | head -3
 timeout 300 python tools/time_train.py 2048 8192 2 128 2>&1 | head -4) 2>&1 | tee gpurun_out/r2v_time.log
