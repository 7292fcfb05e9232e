#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2_c.log 2>&1; tail -4 gpurun_out/pytest_r2_c.log

timeout 900 python bench.py > gpurun_out/bench_r2_c.json 2> gpurun_out/bench_r2_c.err; tail -3 gpurun_out/bench_r2_c.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_c.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
e=d.get("e2e_estimator"); print("est", e and (e["value"], e["ms_per_step"], e["c_abi_variants"]["value"]))
for k,v in d.get("workloads",{}).items():
    e=v.get("e2e_estimator")
    print(k, v.get("value"), v.get("e2e"), v.get("roofline",{}).get("frac"), v.get("max_abs_diff_vs_cpu"), e and (e["value"], e["c_abi_variants"]["value"]), v.get("error"))
PY
