#!/bin/bash
# round 2, call G: 1-GPU suite (incl. the single-rank streamed stripe test), smoke, bench at N=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2g_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
tail -4 gpurun_out/r2g_pytest.log; tail -2 gpurun_out/r2g_smoke.log; tail -5 gpurun_out/r2g_bench.err
