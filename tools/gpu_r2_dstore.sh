#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -m gpu -x -q -k "dm or tma or pipelined or mixed or reset or noise" > gpurun_out/pytest_r2_dstore.log 2>&1; tail -4 gpurun_out/pytest_r2_dstore.log
{
python tools/ab_bench.py
BWQ_FLAGS=2 python tools/ab_bench.py
BWQ_FLAGS=16 python tools/ab_bench.py
} > gpurun_out/ab_dstore.log 2>&1
cat gpurun_out/ab_dstore.log
timeout 900 python bench.py --no-sub-workloads --no-cpu-baseline > gpurun_out/bench_r2_dstore.json 2> gpurun_out/bench_r2_dstore.err; tail -2 gpurun_out/bench_r2_dstore.err
timeout 900 python bench.py --no-sub-workloads --no-cpu-baseline --flags 2 > gpurun_out/bench_r2_nodstore.json 2>> gpurun_out/bench_r2_dstore.err
timeout 900 python bench.py --no-sub-workloads --no-cpu-baseline --workload tfim14_dm --steps 2 --warmup 1 > gpurun_out/bench_r2_dstore_tfim14.json 2>> gpurun_out/bench_r2_dstore.err
python - <<'PY'
import json
for f in ("bench_r2_dstore","bench_r2_nodstore","bench_r2_dstore_tfim14"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["roofline"]["frac"])
    except Exception as e: print(f, "failed", e)
PY
