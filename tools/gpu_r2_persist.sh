#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2_persist.log 2>&1; tail -4 gpurun_out/pytest_r2_persist.log
{
python tools/ab_bench.py
BWQ_FLAGS=32 python tools/ab_bench.py
BWQ_FLAGS=16 python tools/ab_bench.py
} > gpurun_out/ab_persist.log 2>&1
cat gpurun_out/ab_persist.log
timeout 900 python bench.py --no-sub-workloads --no-cpu-baseline > gpurun_out/bench_r2_persist.json 2> gpurun_out/bench_r2_persist.err; tail -2 gpurun_out/bench_r2_persist.err
timeout 900 python bench.py --no-sub-workloads --no-cpu-baseline --flags 32 > gpurun_out/bench_r2_nopersist.json 2> gpurun_out/bench_r2_nopersist.err
timeout 900 python bench.py --no-sub-workloads --no-cpu-baseline --workload tfim14_dm --steps 2 --warmup 1 > gpurun_out/bench_r2_persist_tfim14.json 2>> gpurun_out/bench_r2_persist.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_sweep_tma_persistent -s 4 -c 2 -f -o gpurun_out/tmap_brick10 \
   python tools/profile_case.py brick 10 > gpurun_out/ncu_tmap_brick.log 2>&1
python - <<'PY'
import json
for f in ("bench_r2_persist","bench_r2_nopersist","bench_r2_persist_tfim14"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["roofline"]["frac"])
    except Exception as e: print(f, "failed", e)
PY
