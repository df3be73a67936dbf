#!/bin/bash
# round 2, call S: evidence after the tcgen05 training step: whole 1-GPU suite, smoke, bench at N=1, launch list of the bench
# command, full ncu capture of the tcgen05 training kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2s_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
tail -4 gpurun_out/r2s_pytest.log; tail -2 gpurun_out/r2s_smoke.log; tail -5 gpurun_out/r2s_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --encode-epochs 1"
$CMD > gpurun_out/r2s_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s_launches.csv $CMD > gpurun_out/r2s_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_decode.py train 1024 > gpurun_out/r2s_plain_train.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_fp32 -c 1 -f -o /tmp/prof_train \
    python tools/prof_decode.py train 1024 > gpurun_out/r2s_ncu_train.log 2>&1
echo "train capture rc=$?"
ncu -i /tmp/prof_train.ncu-rep --page raw --csv > gpurun_out/r2s_train_raw.csv 2>/dev/null
ncu -i /tmp/prof_train.ncu-rep --page source --csv > gpurun_out/r2s_train_source.csv 2>/dev/null
ls -la gpurun_out | tail -12
