#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "sv or svx or sharded or statevector" > gpurun_out/pytest_r2_sv3.log 2>&1; tail -4 gpurun_out/pytest_r2_sv3.log
python tools/sv_bench.py 20 26 28 > gpurun_out/sv_bench_r2c.log 2>&1; cat gpurun_out/sv_bench_r2c.log
BWQ_LIB=tools/_build/libbwq_sv3.so python tools/sv_bench.py 20 26 28 > gpurun_out/sv_bench_r2c_b3.log 2>&1; cat gpurun_out/sv_bench_r2c_b3.log
timeout 600 python bench.py --workload tfim30_sv --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2_sv30c.json 2> gpurun_out/bench_r2_sv30c.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_sv30c.json").read().strip().splitlines()[-1])
print("sv30", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("sweep_share_of_step"))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sv_sweep -s 10 -c 2 -f -o gpurun_out/sv_sweep_tfim26_r2d \
   python tools/sv_bench.py 26 > gpurun_out/ncu_sv_r2d.log 2>&1; tail -2 gpurun_out/ncu_sv_r2d.log
