#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout 120 -p no:cacheprovider -k "selftest" > gpurun_out/tc_selftest.log 2>&1; echo "selftest rc=$?"; tail -30 gpurun_out/tc_selftest.log
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout 300 -p no:cacheprovider -s > gpurun_out/tc_decode.log 2>&1; echo "decode tests rc=$?"; tail -40 gpurun_out/tc_decode.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-encode > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"; cat gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
