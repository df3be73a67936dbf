#!/bin/bash
mkdir -p gpurun_out
LBDRN_DEBUG=1 timeout 120 python tools/dbg_wide.py 64 64 > gpurun_out/dbg_notrap.log 2>&1; echo "notrap rc=$?"; tail -8 gpurun_out/dbg_notrap.log
LBDRN_DEBUG=1 timeout 120 python tools/dbg_wide.py 8 16 128 2 > gpurun_out/dbg_notrap2.log 2>&1; echo "notrap rc=$?"; tail -8 gpurun_out/dbg_notrap2.log
