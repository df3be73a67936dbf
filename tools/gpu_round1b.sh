#!/bin/bash
# round-1 (session 2) evidence pass on ONE B200: full GPU test-suite, bench, launch list, ncu captures of both tensor kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_all.log; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1b.json 2> gpurun_out/bench_ref_r1b.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --encode-epochs 1"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_decode.py decode 8192 auto > gpurun_out/plain_decode.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_decode_kernel -s 2 -c 1 -f -o /tmp/prof_decode \
    python tools/prof_decode.py decode 8192 auto > gpurun_out/ncu_decode.log 2>&1
echo "decode capture rc=$?"
ncu -i /tmp/prof_decode.ncu-rep --page raw --csv > gpurun_out/decode_raw.csv 2>/dev/null
ncu -i /tmp/prof_decode.ncu-rep --page source --csv > gpurun_out/decode_source.csv 2>/dev/null
bash tools/gpu_prof_wide.sh
