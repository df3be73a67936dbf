#!/bin/bash
# round 2, call E: decode tests + parity table + timings after the packed-fp32 (FFMA2) epilogues
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_fullsize.py tests/test_gpu_wide.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
timeout 300 python tools/sine_parity.py 1024 2048 5,8 > gpurun_out/r2e_sine_parity.log 2>&1
{
  for p in tensor_fastsin2 tensor_fastsin tensor; do
    echo "pipe"; timeout 120 python tools/time_decode.py 8192 $p 10
  done
  echo "NOPIPE"; LBDRN_TC_NOPIPE=1 timeout 120 python tools/time_decode.py 8192 tensor_fastsin2 10
  echo "wide"; timeout 120 python tools/time_decode.py 8192 auto 3 3 256
} > gpurun_out/r2e_time_decode.log 2>&1
tail -3 gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_sine_parity.log; cat gpurun_out/r2e_time_decode.log
