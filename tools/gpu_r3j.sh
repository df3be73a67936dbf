#!/bin/bash
# round 2, call 3j: role profile of the wide decode kernel (config 3) as it stands
mkdir -p gpurun_out
(LBDRN_TCW_PROF=1 timeout 300 python tools/time_decode.py 4096 auto 2 3 256 2>&1 | grep "lbdrn\|Mpix" | head -12
 timeout 300 python tools/time_decode.py 8192 auto 5 3 256 2>&1 | tail -1) 2>&1 | tee gpurun_out/r3j_tcw_prof.log
