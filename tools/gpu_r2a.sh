#!/bin/bash
# round 2, call A: whole GPU suite, sine-variant parity table, decode kernel A/B timings, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 300 python tools/sine_parity.py 1024 2048 > gpurun_out/r2a_sine_parity.log 2>&1
{
  for p in tensor_fastsin tensor_fastsin2 tensor; do
    for pf in 1 0; do
      echo "PF=$pf"; LBDRN_TC_PF=$pf timeout 120 python tools/time_decode.py 8192 $p 10
    done
  done
  echo "NO_RAW8"; LBDRN_TC_NO_RAW8=1 timeout 120 python tools/time_decode.py 8192 tensor_fastsin 10
  echo "NO_RAW8"; LBDRN_TC_NO_RAW8=1 timeout 120 python tools/time_decode.py 8192 tensor 10
} > gpurun_out/r2a_time_decode.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_sine_parity.log; cat gpurun_out/r2a_time_decode.log
