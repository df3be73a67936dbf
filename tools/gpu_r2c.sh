#!/bin/bash
# round 2, call C: GPU suite, parity table of the pipelined decode kernel, A/B timings (pipelined vs plain two-warpgroup kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
timeout 300 python tools/sine_parity.py 1024 2048 5,8 > gpurun_out/r2c_sine_parity.log 2>&1
{
  for p in tensor_fastsin2 tensor_fastsin tensor; do
    echo "pipe"; timeout 120 python tools/time_decode.py 8192 $p 10
    echo "NOPIPE"; LBDRN_TC_NOPIPE=1 timeout 120 python tools/time_decode.py 8192 $p 10
  done
} > gpurun_out/r2c_time_decode.log 2>&1
tail -5 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_sine_parity.log; cat gpurun_out/r2c_time_decode.log
