#!/bin/bash
# ncu evidence (one GPU): launch list of the bench command + full captures of the top kernels.
# Reports stay in /tmp on the box; CSV exports come back in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --encode-epochs 1"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_decode.py decode 4096 auto > gpurun_out/plain_decode.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_decode_kernel -s 2 -c 1 -f -o /tmp/prof_decode \
    python tools/prof_decode.py decode 4096 auto > gpurun_out/ncu_decode.log 2>&1
echo "decode capture rc=$?"
python tools/prof_decode.py train 1024 > gpurun_out/plain_train.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:train_fp32 -c 1 -f -o /tmp/prof_train \
    python tools/prof_decode.py train 1024 > gpurun_out/ncu_train.log 2>&1
echo "train capture rc=$?"
for n in decode train; do
  ncu -i /tmp/prof_$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$n.ncu-rep --page source --csv > gpurun_out/${n}_source.csv 2>/dev/null
done
du -sh gpurun_out
