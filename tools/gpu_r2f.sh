#!/bin/bash
# round 2, call F: decode tests, timings, ncu capture of the pipelined kernel (mbarrier hand-off instead of warpgroup barriers)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
{
  for p in tensor_fastsin2 tensor_fastsin; do
    echo "pipe"; timeout 120 python tools/time_decode.py 8192 $p 10
  done
} > gpurun_out/r2f_time_decode.log 2>&1
timeout 300 python tools/sine_parity.py 1024 2048 5 > gpurun_out/r2f_sine_parity.log 2>&1
tail -3 gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_time_decode.log gpurun_out/r2f_sine_parity.log
bash tools/gpu_prof_decode.sh r2f
