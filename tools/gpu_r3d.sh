#!/bin/bash
# round 2: tcgen05 bc-64 step with the feature staging buffer over the hidden-output / dz images (coordinates + colours now fit)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/r3d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3d_pytest.log
tail -5 gpurun_out/r3d_pytest.log
(timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 for w in coords coords_col abs d0; do timeout 300 python tools/time_train_cfg.py 2048 $w 2>&1 | head -1; done) 2>&1 | tee gpurun_out/r3d_time.log
