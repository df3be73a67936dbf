#!/bin/bash
# 2-GPU pass: NCCL parity tests + bench at N=1 and N=2 (launched like the driver does)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_dist.log 2>&1; echo "dist tests rc=$?"; tail -25 gpurun_out/pytest_dist.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; cat gpurun_out/bench_n2.json; tail -8 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"; cat gpurun_out/bench_ref_n2.json
