#!/bin/bash
# A/B of the persistent host pool (default build) vs threads spawned per parallel_for (tools/_build/libbwq_nopool.so), N ranks
N=${1:-2}
for i in 1 2; do for l in "" tools/_build/libbwq_nopool.so; do
BWQ_LIB=$l timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 2959$i bench.py --gpus $N --steps 8 --warmup 3 --no-sub-workloads --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('lib=${l:-default}', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), 'est', round(d['e2e_estimator']['ms_per_step'],1), round(d['e2e_estimator']['c_abi_variants']['ms_per_step'],1))"
done; done
