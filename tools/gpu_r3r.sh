#!/bin/bash
# round 2, final state: whole 1-GPU suite, smoke, bench at N=1 (+ reference arm), launch list of the bench command
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r3r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3r_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r3r_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r3r_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r3r_bench.json 2> gpurun_out/r3r_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3r_bench_ref.json 2> gpurun_out/r3r_bench_ref.err; echo "ref rc=$?"
tail -4 gpurun_out/r3r_pytest.log; tail -2 gpurun_out/r3r_smoke.log; tail -5 gpurun_out/r3r_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --encode-epochs 1"
$CMD > gpurun_out/r3r_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3r_launches.csv $CMD > gpurun_out/r3r_ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 300 python tools/time_decode.py 8192 auto 5 3 256 2>&1 | tail -1
