#!/bin/bash
# sharded tfim30_sv: chunked ZZ-layer fusion (BWQ_SVX_ZZ_CHUNK) vs per-pass merging
mkdir -p gpurun_out
N=${1:-2}
for c in ${CHUNKS:-0 3 4 62}; do
BWQ_SVX_ZZ_CHUNK=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 2957$((c%10)) bench.py --gpus $N --workload tfim30_sv --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/zz_$c.json 2> gpurun_out/zz_$c.err
python - <<PY
import json
try:
    s=json.loads([l for l in open("gpurun_out/zz_$c.json").read().splitlines() if l.startswith("{")][-1])
    e=s["exchange"]; print("chunk=$c N=$N", round(s["value"],2), "circ/s", round(s["ms_per_step"],2), "ms frac", round(s["roofline"]["frac"],3), "exch/circ", e["exchanges_per_circuit"], "exch ms", round(e["ms_per_step"],2), "diff", s.get("max_abs_diff_vs_1rank"))
except Exception as ex:
    print("chunk=$c failed", ex, open("gpurun_out/zz_$c.err").read()[-800:])
PY
done
