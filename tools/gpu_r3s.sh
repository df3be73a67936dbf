#!/bin/bash
# round 2, call 3s: final bench lines (N=1: own arm + reference arm) after the encode leg reports warm and first-run times
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r3s_bench.json 2> gpurun_out/r3s_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3s_bench_ref.json 2> gpurun_out/r3s_bench_ref.err; echo "ref rc=$?"
tail -3 gpurun_out/r3s_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3s_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e']['frac_of_host_copy_ceiling'])
e=d['encode']; print('encode', e['s_per_scene'], e['first_run_s_per_scene'], e['reference_sampler_s_per_scene'])
PY
