#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2_sv2.log 2>&1; tail -6 gpurun_out/pytest_r2_sv2.log
{
python tools/ab_bench.py
BWQ_FLAGS=2 python tools/ab_bench.py
} > gpurun_out/ab_dstore.log 2>&1
cat gpurun_out/ab_dstore.log
timeout 900 python bench.py --no-sub-workloads --no-cpu-baseline > gpurun_out/bench_r2_dstore.json 2> gpurun_out/bench_r2_dstore.err; tail -2 gpurun_out/bench_r2_dstore.err
timeout 900 python bench.py --no-sub-workloads --no-cpu-baseline --flags 2 > gpurun_out/bench_r2_nodstore.json 2>> gpurun_out/bench_r2_dstore.err
timeout 900 python bench.py --no-sub-workloads --no-cpu-baseline --workload tfim14_dm --steps 2 --warmup 1 > gpurun_out/bench_r2_dstore_tfim14.json 2>> gpurun_out/bench_r2_dstore.err
timeout 600 python bench.py --workload tfim30_sv --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2_sv30b.json 2> gpurun_out/bench_r2_sv30b.err
python - <<'PY'
import json
for f in ("bench_r2_dstore","bench_r2_nodstore","bench_r2_dstore_tfim14","bench_r2_sv30b"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"])
    except Exception as e: print(f, "failed", e)
PY
python tools/sv_bench.py 20 24 26 28 > gpurun_out/sv_bench_r2b.log 2>&1; cat gpurun_out/sv_bench_r2b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sv_sweep -s 10 -c 2 -f -o gpurun_out/sv_sweep_tfim26_r2b \
   python tools/sv_bench.py 26 > gpurun_out/ncu_sv_r2b.log 2>&1; tail -2 gpurun_out/ncu_sv_r2b.log
