#!/bin/bash
# round 2, call 3m: compute-sanitizer memcheck over the training tests that exercise the aligned-word gathers (reads next
# to the ends of the plane buffers) -- small scenes, so the 10-50x slowdown stays within minutes
mkdir -p gpurun_out
timeout 1200 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 7 \
  python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "interleaved_chunk or sixteen_bit or odd_scene" \
  > gpurun_out/r3m_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r3m_memcheck.log
grep -c "Invalid\|out of bounds" gpurun_out/r3m_memcheck.log
tail -15 gpurun_out/r3m_memcheck.log
