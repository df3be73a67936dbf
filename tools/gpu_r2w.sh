#!/bin/bash
# round 2, call W (2 GPUs): NCCL tests + bench at N=2 with the session's training / evaluation / sampler changes
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2w_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err; echo "bench rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2w_bench_ref_n2.json 2> gpurun_out/r2w_bench_ref_n2.err; echo "ref rc=$?"
tail -3 gpurun_out/r2w_pytest_dist.log; tail -3 gpurun_out/r2w_bench_n2.err; cut -c1-600 gpurun_out/r2w_bench_n2.json; cut -c1-400 gpurun_out/r2w_bench_ref_n2.json
