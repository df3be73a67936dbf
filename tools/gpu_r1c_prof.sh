#!/bin/bash
# round-1 session-3 evidence: launch list of the bench command + full ncu capture of the fp16-split training kernel
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --encode-epochs 1"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_decode.py train 1024 > gpurun_out/plain_train.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_fp32 -c 1 -f -o /tmp/prof_train \
    python tools/prof_decode.py train 1024 > gpurun_out/ncu_train.log 2>&1
echo "train capture rc=$?"
ncu -i /tmp/prof_train.ncu-rep --page raw --csv > gpurun_out/train_raw.csv 2>/dev/null
ncu -i /tmp/prof_train.ncu-rep --page source --csv > gpurun_out/train_source.csv 2>/dev/null
(timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | head -1
 timeout 300 python tools/time_train.py 2048 8192 2 128 2>&1 | head -1
 LBDRN_TRAIN_TF32=1 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | head -1
 LBDRN_TRAIN_CHW=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1) | tee gpurun_out/time_train_variants.log
