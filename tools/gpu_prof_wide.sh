#!/bin/bash
# full ncu capture of the wide tensor decode kernel (D=3 bc256) on a 2048^2 scene
mkdir -p gpurun_out
python tools/time_decode.py 4096 auto 3 3 256 > gpurun_out/plain_wide.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tcw_decode_kernel -s 2 -c 1 -f -o /tmp/prof_wide \
    python tools/time_decode.py 4096 auto 3 3 256 > gpurun_out/ncu_wide.log 2>&1
echo "wide capture rc=$?"
ncu -i /tmp/prof_wide.ncu-rep --page raw --csv > gpurun_out/wide_raw.csv 2>/dev/null
ncu -i /tmp/prof_wide.ncu-rep --page source --csv > gpurun_out/wide_source.csv 2>/dev/null
ls -la gpurun_out/wide_*.csv
