#!/bin/bash
# round 2, final state at N=8: bench (decode weak scaling, e2e, config 5 on 8 GPUs, strong scaling of one scene, scene-per-GPU and DP encode)
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r3b_bench_n8.json 2> gpurun_out/r3b_bench_n8.err; echo "bench rc=$?"
tail -3 gpurun_out/r3b_bench_n8.err; cut -c1-300 gpurun_out/r3b_bench_n8.json
