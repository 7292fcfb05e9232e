#!/bin/bash
# round 2: TMA sweep kernel -- parity, A/B against the classic kernel and occupancy variants, bench, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2_tma.log 2>&1; tail -4 gpurun_out/pytest_r2_tma.log
{
python tools/ab_bench.py
BWQ_FLAGS=16 python tools/ab_bench.py
BWQ_FLAGS=0xffff00 python tools/ab_bench.py
BWQ_LIB=tools/_build/libbwq_b3.so python tools/ab_bench.py
BWQ_LIB=tools/_build/libbwq_b5.so python tools/ab_bench.py
} > gpurun_out/ab_tma.log 2>&1
cat gpurun_out/ab_tma.log
timeout 900 python bench.py > gpurun_out/bench_r2_tma.json 2> gpurun_out/bench_r2_tma.err; tail -2 gpurun_out/bench_r2_tma.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_sweep_tma -s 6 -c 2 -f -o gpurun_out/tma_brick10 \
   python tools/profile_case.py brick 10 > gpurun_out/ncu_tma_brick.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_sweep_tma -s 13 -c 2 -f -o gpurun_out/tma_tfim13 \
   python tools/profile_case.py tfim 13 > gpurun_out/ncu_tma_tfim.log 2>&1
tail -2 gpurun_out/ncu_tma_brick.log gpurun_out/ncu_tma_tfim.log
