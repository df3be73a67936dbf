#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/pytest_train_full.log 2>&1; echo "train rc=$?"
grep -v "^frame\|^#" gpurun_out/pytest_train_full.log | tail -3
(LBDRN_TRAIN_PROF=1 timeout 600 python tools/time_train.py 8192 8192 2 64 2>&1 | grep -v "^Traceback\|File\|print\|Broken" | head -2
 LBDRN_TRAIN_PROF=1 timeout 600 python tools/time_train.py 2048 8192 2 64 2>&1 | head -2) | tee gpurun_out/time_train.log
python tools/prof_decode.py train 1024 > gpurun_out/plain_train.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_fp32 -c 1 -f -o /tmp/prof_train \
    python tools/prof_decode.py train 1024 > gpurun_out/ncu_train.log 2>&1
echo "train capture rc=$?"
ncu -i /tmp/prof_train.ncu-rep --page raw --csv > gpurun_out/train_raw.csv 2>/dev/null
ncu -i /tmp/prof_train.ncu-rep --page source --csv > gpurun_out/train_source.csv 2>/dev/null
