#!/bin/bash
# ncu evidence: launch list of the bench command + full capture of the top kernels (one GPU).
# Reports stay in /tmp on the box; only CSV exports (and the .ncu-rep when small) come back in gpurun_out/.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --encode-epochs 1 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --encode-epochs 1 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_decode.py decode 2048 > gpurun_out/plain_decode.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"infer_fp32|tc_decode" -s 1 -c 1 -f -o /tmp/prof_decode \
    python tools/prof_decode.py decode 2048 > gpurun_out/ncu_decode.log 2>&1
echo "decode capture rc=$?"
python tools/prof_decode.py train 512 > gpurun_out/plain_train.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:train_fp32 -c 1 -f -o /tmp/prof_train \
    python tools/prof_decode.py train 512 > gpurun_out/ncu_train.log 2>&1
echo "train capture rc=$?"
for n in decode train; do
  ncu -i /tmp/prof_$n.ncu-rep --page raw --csv > gpurun_out/${n}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$n.ncu-rep --page details --csv > gpurun_out/${n}_details.csv 2>/dev/null
  ncu -i /tmp/prof_$n.ncu-rep --page source --csv > gpurun_out/${n}_source.csv 2>/dev/null
  sz=$(stat -c %s /tmp/prof_$n.ncu-rep); echo "$n report $sz bytes"
  if [ "$sz" -lt 12000000 ]; then cp /tmp/prof_$n.ncu-rep gpurun_out/; fi
done
du -sh gpurun_out
