#!/bin/bash
# round 2, call O: A/B of the reduction shape (8 vs 4 row groups) and the weight reload (bulk copies vs cp.async)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q  > gpurun_out/r2o_pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest_train.log
tail -4 gpurun_out/r2o_pytest_train.log
(for v in ""; do
 echo "== $v"
 env $v LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep "train phases" | head -1 | cut -c1-800
 env $v timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 done
 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | head -1
 LBDRN_TRAIN_H2=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1) 2>&1 | tee gpurun_out/r2o_time_train.log
