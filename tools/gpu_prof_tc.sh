#!/bin/bash
# full ncu capture of the main tensor decode kernel only
mkdir -p gpurun_out
python tools/prof_decode.py decode 4096 auto > gpurun_out/plain_decode.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_decode_kernel -s 2 -c 1 -f -o /tmp/prof_decode \
    python tools/prof_decode.py decode 4096 auto > gpurun_out/ncu_decode.log 2>&1
echo "decode capture rc=$?"
ncu -i /tmp/prof_decode.ncu-rep --page raw --csv > gpurun_out/decode_raw.csv 2>/dev/null
ncu -i /tmp/prof_decode.ncu-rep --page source --csv > gpurun_out/decode_source.csv 2>/dev/null
