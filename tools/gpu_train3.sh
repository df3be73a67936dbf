#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/pytest_train_full.log 2>&1; echo "train rc=$?"
grep -v "^frame\|^#" gpurun_out/pytest_train_full.log | tail -6
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_enc.json 2> gpurun_out/bench_enc.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_enc.json').read().strip().splitlines()[-1]); print(d['value'], d['encode'])"
