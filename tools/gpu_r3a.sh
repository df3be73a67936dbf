#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "other_shapes" > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3a_pytest.log
tail -30 gpurun_out/r3a_pytest.log
