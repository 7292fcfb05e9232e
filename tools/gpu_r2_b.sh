#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2_b.log 2>&1; tail -4 gpurun_out/pytest_r2_b.log
python tools/sv_bench.py 20 26 28 > gpurun_out/sv_bench_r2d.log 2>&1; cat gpurun_out/sv_bench_r2d.log
timeout 900 python bench.py > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; tail -3 gpurun_out/bench_r2_b.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_b.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
e=d.get("e2e_estimator"); print("est", e and (e["value"], e["ms_per_step"], e["c_abi_variants"]))
for k,v in d.get("workloads",{}).items():
    e=v.get("e2e_estimator")
    print(k, v.get("value"), v.get("e2e"), v.get("roofline",{}).get("frac"), v.get("max_abs_diff_vs_cpu"), e and (e["value"], e["c_abi_variants"]["value"]), v.get("error"))
PY
