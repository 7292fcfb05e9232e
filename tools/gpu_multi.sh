#!/bin/bash
# N-GPU checks: sharded statevector parity at world 2/4/8, cfg4 bench, cfg2 weak-scaling bench.
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --workload tfim30_sv --steps 4 --warmup 2 2>/dev/null | tail -1 > gpurun_out/bench_tfim30sv_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_brick10_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload tfim14_dm --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_tfim14_n$N.json
for f in gpurun_out/bench_tfim30sv_n$N.json gpurun_out/bench_brick10_n$N.json gpurun_out/bench_tfim14_n$N.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read()); r=d["roofline"]
print(sys.argv[1], {k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], "frac", r["frac"], "exchange", d.get("exchange"))
PY
done
