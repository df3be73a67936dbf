#!/bin/bash
# round 2, call I (2 GPUs): streamed-stripe test + bench at N=2 after the steady-state single-call change; training timings on GPU 0
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2i_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-encode > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "bench rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/time_train.py 8192 > gpurun_out/r2i_time_train.log 2>&1
tail -3 gpurun_out/r2i_pytest_dist.log; tail -3 gpurun_out/r2i_bench_n2.err; cat gpurun_out/r2i_time_train.log
