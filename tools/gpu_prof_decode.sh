#!/bin/bash
# full ncu capture of the decode kernel only (one GPU); CSV exports come back in gpurun_out/
mkdir -p gpurun_out
SIDE=${1:-2048}; PATHSEL=${2:-auto}
python tools/prof_decode.py decode $SIDE $PATHSEL > gpurun_out/plain_decode.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${3:-infer_fp32|tc_decode}" -s 1 -c 1 -f -o /tmp/prof_decode \
    python tools/prof_decode.py decode $SIDE $PATHSEL > gpurun_out/ncu_decode.log 2>&1
echo "decode capture rc=$?"
ncu -i /tmp/prof_decode.ncu-rep --page raw --csv > gpurun_out/decode_raw.csv 2>/dev/null
ncu -i /tmp/prof_decode.ncu-rep --page details --csv > gpurun_out/decode_details.csv 2>/dev/null
ncu -i /tmp/prof_decode.ncu-rep --page source --csv > gpurun_out/decode_source.csv 2>/dev/null
tail -2 gpurun_out/plain_decode.log gpurun_out/ncu_decode.log; du -sh gpurun_out
