#!/bin/bash
# round 2, call 3q: are the encode lines of the r3p bench a slow box or a regression?  per-step timings of the feature-set variants + the bench encode legs again
mkdir -p gpurun_out
(for w in coords coords_col abs d0 c8; do timeout 300 python tools/time_train_cfg.py 8192 $w 2>&1 | head -1; done
 timeout 300 python tools/time_train.py 8192 8192 2 64 2>&1 | head -1
 LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train_cfg.py 8192 coords 2>&1 | grep "train phases" | head -1 | cut -c1-900) 2>&1 | tee gpurun_out/r3q_time.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3q_bench.json 2> gpurun_out/r3q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3q_bench.json').read().strip().splitlines()[-1])
print('encode', d['encode']['s_per_scene'], d['encode']['reference_sampler_s_per_scene'])
for v in d['other_configs']:
    if 'ENCODE' in v['config']: print(v['config'][:30], v.get('s_per_scene_10_epochs_extrapolated'), v.get('us_per_step_incl_eval'))
PY
