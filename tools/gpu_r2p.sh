#!/bin/bash
# round 2, call P: MUFU sine/cosine in the tcgen05 training step: parity of the whole training suite with it on, and the gain
mkdir -p gpurun_out
LBDRN_TRAIN_FASTSIN=1 timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/r2p_pytest_fastsin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest_fastsin.log
tail -12 gpurun_out/r2p_pytest_fastsin.log
(for v in "" "LBDRN_TRAIN_FASTSIN=1"; do
 echo "== $v"
 env $v LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep "train phases" | head -1 | cut -c1-800
 env $v timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 done) 2>&1 | tee gpurun_out/r2p_time_train.log
