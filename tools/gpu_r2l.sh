#!/bin/bash
# round 2, call L: tcgen05 training step with thread-block clusters (DSMEM pre-reduction, multicast weight reload)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "gradients or trajectory or odd_scene or determin or bit_identical" > gpurun_out/r2l_pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_train.log
tail -5 gpurun_out/r2l_pytest_train.log
(for cl in 4 2 1 8; do
 LBDRN_TRAIN_CLUSTER=$cl LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep "train phases" | head -1 | cut -c1-700
 LBDRN_TRAIN_CLUSTER=$cl timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 done
 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | head -1) 2>&1 | tee gpurun_out/r2l_time_train.log
