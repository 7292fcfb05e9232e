#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2_d.log 2>&1; tail -4 gpurun_out/pytest_r2_d.log
timeout 600 python bench.py --workload tfim4_lima_zne --steps 5 --warmup 3 --no-sub-workloads > gpurun_out/bench_r2_cfg1.json 2> gpurun_out/bench_r2_cfg1.err; tail -5 gpurun_out/bench_r2_cfg1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_cfg1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"])
print("roofline", d["roofline"])
print("est", d.get("e2e_estimator"))
print("cpu", d.get("cpu_baseline"))
PY
