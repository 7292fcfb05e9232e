#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2_var.log 2>&1; tail -6 gpurun_out/pytest_r2_var.log
timeout 900 python bench.py > gpurun_out/bench_r2_var.json 2> gpurun_out/bench_r2_var.err; tail -3 gpurun_out/bench_r2_var.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_var.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "est", d.get("e2e_estimator"))
for k,v in d.get("workloads",{}).items(): print(k, v.get("value"), v.get("e2e"), v.get("roofline",{}).get("frac"), v.get("max_abs_diff_vs_cpu"), v.get("e2e_estimator",{}).get("value") if v.get("e2e_estimator") else None, v.get("error"))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sv_sweep -s 10 -c 2 -f -o gpurun_out/sv_sweep_tfim26_r2c \
   python tools/sv_bench.py 26 > gpurun_out/ncu_sv_r2c.log 2>&1; tail -3 gpurun_out/ncu_sv_r2c.log
