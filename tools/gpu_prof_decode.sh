#!/bin/bash
# ncu --set full capture of the headline decode kernel at 8192^2 (after a plain run of the same command exited 0)
# usage: gpu_prof_decode.sh <tag> [kernel regex]
tag=${1:-r2}; re=${2:-tc_pipe_kernel}
mkdir -p gpurun_out
python tools/prof_decode.py decode 8192 auto > gpurun_out/${tag}_plain_decode.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$re -s 2 -c 1 -f -o /tmp/prof_decode \
    python tools/prof_decode.py decode 8192 auto > gpurun_out/${tag}_ncu_decode.log 2>&1
echo "decode capture rc=$?"
ncu -i /tmp/prof_decode.ncu-rep --page raw --csv > gpurun_out/${tag}_decode_raw.csv 2>/dev/null
ncu -i /tmp/prof_decode.ncu-rep --page source --csv > gpurun_out/${tag}_decode_source.csv 2>/dev/null
