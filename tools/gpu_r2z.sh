#!/bin/bash
# round 2, call Z: final evidence of the session: whole 1-GPU suite, smoke, bench at N=1, launch list of the bench command,
# full ncu captures of the two tcgen05 training kernels (bc 64 resident, bc 256 streamed) and of the wide evaluation kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2z_smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; echo "ref rc=$?"
tail -4 gpurun_out/r2z_pytest.log; tail -2 gpurun_out/r2z_smoke.log; tail -5 gpurun_out/r2z_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --encode-epochs 1"
$CMD > gpurun_out/r2z_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv $CMD > gpurun_out/r2z_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_decode.py train 1024 > gpurun_out/r2z_plain_train.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_fp32 -c 1 -f -o /tmp/prof_train \
    python tools/prof_decode.py train 1024 > gpurun_out/r2z_ncu_train.log 2>&1
echo "train capture rc=$?"
ncu -i /tmp/prof_train.ncu-rep --page raw --csv > gpurun_out/r2z_train_raw.csv 2>/dev/null
ncu -i /tmp/prof_train.ncu-rep --page source --csv > gpurun_out/r2z_train_source.csv 2>/dev/null
python tools/prof_train256.py 1024 > gpurun_out/r2z_plain_train256.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:train_fp32 -c 1 -f -o /tmp/prof_train256 \
    python tools/prof_train256.py 1024 > gpurun_out/r2z_ncu_train256.log 2>&1
echo "train256 capture rc=$?"
ncu -i /tmp/prof_train256.ncu-rep --page raw --csv > gpurun_out/r2z_train256_raw.csv 2>/dev/null
ncu -i /tmp/prof_train256.ncu-rep --page source --csv > gpurun_out/r2z_train256_source.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tcw_decode_kernel -c 1 -f -o /tmp/prof_eval256 \
    python tools/prof_train256.py 1024 > gpurun_out/r2z_ncu_eval256.log 2>&1
echo "eval256 capture rc=$?"
ncu -i /tmp/prof_eval256.ncu-rep --page raw --csv > gpurun_out/r2z_eval256_raw.csv 2>/dev/null
ncu -i /tmp/prof_eval256.ncu-rep --page source --csv > gpurun_out/r2z_eval256_source.csv 2>/dev/null
ls -la gpurun_out | grep r2z | awk '{print $5, $9}'
