#!/bin/bash
# Round measurement set: benches of every BASELINE config that fits one GPU, launch list, full ncu captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python bench.py > gpurun_out/bench_brick10.json 2> gpurun_out/bench_brick10.err; echo "brick10 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
timeout 600 python bench.py --workload tfim4_lima_zne > gpurun_out/bench_tfim4.json 2> gpurun_out/bench_tfim4.err; echo "tfim4 rc=$?"
timeout 900 python bench.py --workload tfim14_dm --steps 2 --warmup 1 > gpurun_out/bench_tfim14.json 2> gpurun_out/bench_tfim14.err; echo "tfim14 rc=$?"
timeout 600 python bench.py --workload tfim12_dm --steps 3 --warmup 1 --scale 4 > gpurun_out/bench_tfim12.json 2> gpurun_out/bench_tfim12.err; echo "tfim12 rc=$?"
timeout 600 python bench.py --workload tfim30_sv --steps 3 --warmup 1 > gpurun_out/bench_tfim30sv.json 2> gpurun_out/bench_tfim30sv.err; echo "tfim30sv rc=$?"
timeout 900 python bench.py --workload mixed6_12_dataset --steps 2 --warmup 1 > gpurun_out/bench_mixed6_12.json 2> gpurun_out/bench_mixed6_12.err; echo "mixed rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --scale 0.1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_sweep -s 6 -c 2 -f -o gpurun_out/sweep_brick10 \
   python tools/profile_case.py brick 10 > gpurun_out/ncu_full_brick.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_sweep -s 13 -c 2 -f -o gpurun_out/sweep_tfim13 \
   python tools/profile_case.py tfim 13 > gpurun_out/ncu_full_tfim.log 2>&1
for f in gpurun_out/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "frac", r.get("frac"), "GB/s", r.get("achieved"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("sample"), "diff", (d.get("cpu_baseline") or {}).get("max_abs_diff_vs_gpu"))
except Exception as e:
    print("parse failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
