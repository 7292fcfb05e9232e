#!/bin/bash
# One gpurun call: parity tests, smoke, bench, launch list, one full ncu capture of the sweep kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python tools/quick_bench.py > gpurun_out/quick_bench.log 2>&1
cat gpurun_out/quick_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scale 0.1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_sweep -s 20 -c 3 -f -o gpurun_out/sweep_brick10 \
   python tools/profile_case.py brick 10 > gpurun_out/ncu_full_brick.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_sweep -s 6 -c 3 -f -o gpurun_out/sweep_tfim12 \
   python tools/profile_case.py tfim 12 > gpurun_out/ncu_full_tfim.log 2>&1
ls -la gpurun_out
