#!/bin/bash
# 2-GPU checks: sharded statevector parity test, the driver's multi-GPU bench launch, reference arm under torchrun
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus" 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -3 gpurun_out/bench_r2_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2_n2_ref.json 2> gpurun_out/bench_r2_n2_ref.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --workload tfim30_sv --steps 4 --warmup 1 > gpurun_out/bench_r2_n2_sv30.json 2> gpurun_out/bench_r2_n2_sv30.err
python - <<'PY'
import json
def last(f):
    try: return json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    except Exception as e: return {"error": repr(e)}
d=last("gpurun_out/bench_r2_n2.json")
print("n2", d.get("value"), d.get("e2e",{}).get("value"), d.get("roofline",{}).get("frac"), d.get("error"))
for k,v in d.get("workloads",{}).items():
    print("  ", k, v.get("value"), v.get("roofline",{}).get("frac"), v.get("exchange"), v.get("max_abs_diff_vs_1rank"), v.get("error"))
r=last("gpurun_out/bench_r2_n2_ref.json"); print("ref", r.get("value"), r.get("cpu_baseline"), r.get("config"))
s=last("gpurun_out/bench_r2_n2_sv30.json"); print("sv30 n2", s.get("value"), s.get("ms_per_step"), s.get("roofline",{}).get("frac"), s.get("exchange"), s.get("max_abs_diff_vs_1rank"))
PY
