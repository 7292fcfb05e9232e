#!/bin/bash
# what the driver runs at round end, on one box: pytest -m gpu, smoke(), bench.py (+ reference arm)
mkdir -p gpurun_out/verify
O=gpurun_out/verify
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log | cut -c1-300
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/verify/bench_default.json").read().strip().splitlines()[-1])
print("lines on stdout:", len(open("gpurun_out/verify/bench_default.json").read().strip().splitlines()))
print({k:d[k] for k in ("metric","value","unit","n_gpus","steps","warmup","ms_per_step","higher_is_better","scaling","vs_baseline","dtype","data","gpu_launches")})
print("e2e", d["e2e"]); print("roofline", {k:d["roofline"][k] for k in ("bound","achieved","peak","unit","frac","traffic","kernel")}); print("cpu", d["cpu_baseline"]); print("clocks", d.get("clocks"))
print("est", d["e2e_estimator"]["value"], d["e2e_estimator"]["c_abi_variants"]["value"])
for k,v in d["workloads"].items(): print(k, round(v["value"],2), round(v["e2e"],2), round(v["roofline"]["frac"],4), v["roofline"]["kernel"], v.get("max_abs_diff_vs_cpu"), (v.get("e2e_estimator") or {}).get("c_abi_variants",{}).get("value"), v.get("error"))
r=json.loads(open("gpurun_out/verify/bench_reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["cpu_baseline"]["cores"], r["impl"])
PY
