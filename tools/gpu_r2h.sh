#!/bin/bash
# round 2, call H (2 GPUs): NCCL parity tests + bench at N=2 (+ the CLI test of the built-in nn codec)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py "tests/test_gpu_train.py::test_cli_roundtrip_with_the_builtin_nn_codec" -m gpu -x -q > gpurun_out/r2h_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "bench rc=$?"
tail -4 gpurun_out/r2h_pytest_dist.log; tail -5 gpurun_out/r2h_bench_n2.err
