#!/bin/bash
# round 2, call T: training step on the other feature sets; encode wall with the native reference sampler
mkdir -p gpurun_out
(for w in coords coords_col abs d0 c8; do LBDRN_TRAIN_PROF=1 timeout 300 python tools/time_train_cfg.py 2048 $w 2>&1 | grep -v "^$" | cut -c1-900 | head -4; done
 timeout 600 python tools/enc_gap.py 2>&1 | tail -6) 2>&1 | tee gpurun_out/r2t_cfg.log
