#!/bin/bash
# 8-GPU: the driver's multi-GPU bench launch (default workloads) + sharded tfim30_sv alone
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; tail -3 gpurun_out/bench_r2_n$N.err
python - <<PY
import json
def last(f):
    try: return json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    except Exception as e: return {"error": repr(e), "tail": open(f.replace(".json",".err")).read()[-2500:]}
d=last("gpurun_out/bench_r2_n$N.json")
print("n$N", d.get("value"), d.get("e2e",{}).get("value"), d.get("roofline",{}).get("frac"), d.get("error"), d.get("tail"))
for k,v in d.get("workloads",{}).items():
    print("  ", k, v.get("value"), v.get("ms_per_step"), v.get("roofline",{}).get("frac"), v.get("exchange"), v.get("max_abs_diff_vs_1rank"), v.get("error"))
PY
