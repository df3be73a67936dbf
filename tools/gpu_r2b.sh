#!/bin/bash
# round 2, call B: GPU suite (after fixes) + ncu --set full capture of the headline decode kernel at 8192^2
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
python tools/prof_decode.py decode 8192 auto > gpurun_out/r2b_plain_decode.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_decode_kernel -s 2 -c 1 -f -o /tmp/prof_decode \
    python tools/prof_decode.py decode 8192 auto > gpurun_out/r2b_ncu_decode.log 2>&1
echo "decode capture rc=$?"
ncu -i /tmp/prof_decode.ncu-rep --page raw --csv > gpurun_out/r2b_decode_raw.csv 2>/dev/null
ncu -i /tmp/prof_decode.ncu-rep --page source --csv > gpurun_out/r2b_decode_source.csv 2>/dev/null
tail -4 gpurun_out/r2b_pytest.log
