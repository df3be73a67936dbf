#!/bin/bash
# round 2, call 3n: ncu --set full capture of the bc 256 streamed tcgen05 training kernel in its final state (after the chunk-gather change)
mkdir -p gpurun_out
python tools/prof_train256.py 1024 > gpurun_out/r3n_plain_train256.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:train_fp32 -c 1 -f -o /tmp/prof_train256 \
    python tools/prof_train256.py 1024 > gpurun_out/r3n_ncu_train256.log 2>&1
echo "train256 capture rc=$?"
ncu -i /tmp/prof_train256.ncu-rep --page raw --csv > gpurun_out/r3n_train256_raw.csv 2>/dev/null
ncu -i /tmp/prof_train256.ncu-rep --page source --csv > gpurun_out/r3n_train256_source.csv 2>/dev/null
ls -la gpurun_out | grep r3n | awk '{print $5, $9}'
