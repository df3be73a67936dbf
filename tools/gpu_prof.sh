#!/bin/bash
# ncu evidence: launch list of the bench command + full capture of the top kernels (one GPU)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --encode-epochs 1 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --encode-epochs 1 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_decode.py decode 2048 > gpurun_out/plain_decode.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:infer_fp32 -s 1 -c 1 -f -o gpurun_out/prof_decode \
    python tools/prof_decode.py decode 2048 > gpurun_out/ncu_decode.log 2>&1
echo "decode capture rc=$?"
python tools/prof_decode.py train 1024 > gpurun_out/plain_train.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:train_fp32 -c 1 -f -o gpurun_out/prof_train \
    python tools/prof_decode.py train 1024 > gpurun_out/ncu_train.log 2>&1
echo "train capture rc=$?"
tail -3 gpurun_out/plain_decode.log gpurun_out/plain_train.log gpurun_out/ncu_decode.log gpurun_out/ncu_train.log
