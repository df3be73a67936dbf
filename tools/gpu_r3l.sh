#!/bin/bash
# round 2, call 3l (2 GPUs): distributed suite + N=2 bench on the final tree (chunk-gather change touches the data-parallel trainer)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r3l_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3l_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-encode > gpurun_out/r3l_bench_n2.json 2> gpurun_out/r3l_bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/r3l_pytest_dist.log; tail -3 gpurun_out/r3l_bench_n2.err
