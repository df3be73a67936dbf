#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/pytest_train_full.log 2>&1; echo "train rc=$?"
grep -v "^frame\|^#" gpurun_out/pytest_train_full.log | head -60
