#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_enc$i.json 2> gpurun_out/bench_enc$i.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_enc$i.json'))
print({k:d['encode'][k] for k in ('s_per_scene','us_per_step_incl_eval','final_val_mse')}, d['value'], d['encode']['clocks'])
PY
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv
lscpu | grep -E "Model name|^CPU\(s\)"
