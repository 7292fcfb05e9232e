#!/bin/bash
# 2-GPU: sharded statevector with the exchange fused into the sweeps vs separate exchange kernels
mkdir -p gpurun_out
nvidia-smi -L | wc -l; N=${1:-2}
[ "$N" = 2 ] && timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus" 2>&1 | tail -3
for f in 1 0; do
BWQ_SVX_FUSED_EXCHANGE=$f timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 2954$f bench.py --gpus $N --workload tfim30_sv --steps 4 --warmup 1 > gpurun_out/bench_r2_n2_sv30_n${N}_fused$f.json 2> gpurun_out/bench_r2_n2_sv30_n${N}_fused$f.err
python - <<PY
import json
def last(f):
    try: return json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    except Exception as e: return {"error": repr(e), "tail": open(f.replace(".json",".err")).read()[-1500:]}
s=last("gpurun_out/bench_r2_n2_sv30_n${N}_fused$f.json"); print("fused=$f sv30 n2", s.get("value"), s.get("ms_per_step"), s.get("roofline",{}).get("frac"), s.get("exchange"), s.get("max_abs_diff_vs_1rank"), s.get("error"), s.get("tail"))
PY
done
