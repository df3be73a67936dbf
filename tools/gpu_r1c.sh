#!/bin/bash
# round-1 session-3 GPU pass: whole GPU suite, smoke, bench (own arm + reference arm), training timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"
grep -v "^frame\|^#" gpurun_out/pytest_all.log | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_r1c.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1c.json 2> gpurun_out/bench_ref_r1c.err; echo "ref rc=$?"
(LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | grep -v "^Traceback\|File\|print\|Broken" | head -3 | cut -c1-600
 timeout 300 python tools/time_train.py 2048 8192 2 64 2>&1 | head -1) | tee gpurun_out/time_train.log
