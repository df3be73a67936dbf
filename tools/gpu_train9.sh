#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/pytest_train_full.log 2>&1; echo "train rc=$?"
grep -v "^frame\|^#" gpurun_out/pytest_train_full.log | tail -3
timeout 500 python tools/enc_gap.py 2>&1 | tail -4 | tee gpurun_out/enc_gap.log
