#!/bin/bash
# tests + bench + quick benches (no ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"frac",round(d["roofline"]["frac"],3),"GB/s",round(d["roofline"]["achieved"]),"cpu",d["cpu_baseline"] and round(d["cpu_baseline"]["value"],1), "err", d["cpu_baseline"] and d["cpu_baseline"]["max_abs_diff_vs_gpu"])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
timeout 300 python tools/quick_bench.py 2>&1 | tee gpurun_out/quick_bench.log
timeout 300 python tools/sv_bench.py 16 24 28 2>&1 | tee gpurun_out/sv_bench.log
