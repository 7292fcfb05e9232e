#!/bin/bash
# Round-2 measurement set (1 GPU): tests, the driver's bench line + reference arm, launch list of the bench
# command, ncu --set full captures of the three hot kernels, sanitizer pass.
mkdir -p gpurun_out/final
O=gpurun_out/final
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > $O/gpu.txt; nproc >> $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
timeout 600 python bench.py --workload mixed6_12_dataset --steps 2 --warmup 1 --no-sub-workloads > $O/bench_mixed6_12.json 2> $O/bench_mixed6_12.err; echo "mixed rc=$?"
timeout 600 python bench.py --workload tfim12_dm --steps 3 --warmup 1 --scale 4 --no-sub-workloads > $O/bench_tfim12.json 2> $O/bench_tfim12.err; echo "tfim12 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_default.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-workloads > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_sweep_tma -s 6 -c 2 -f -o $O/tma_brick10 \
   python tools/profile_case.py brick 10 > $O/ncu_tma_brick.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_sweep_tma -s 13 -c 2 -f -o $O/tma_tfim13 \
   python tools/profile_case.py tfim 13 > $O/ncu_tma_tfim.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sv_sweep_kernel -s 20 -c 1 -f -o $O/sv_sweep_tfim26 \
   python tools/sv_bench.py 26 > $O/ncu_sv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_onchip_kernel -s 19 -c 1 -f -o $O/onchip_cfg1 \
   python tools/cfg1_breakdown.py > $O/ncu_onchip.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_onchip_gpu.py -m gpu -x -q -k "all_gates or status" > $O/sanitizer_memcheck_onchip.log 2>&1; tail -3 $O/sanitizer_memcheck_onchip.log
timeout 300 python tools/cfg1_breakdown.py 2>&1 | grep -v host_threads > $O/cfg1_breakdown.log
for f in $O/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "e2e", (d.get("e2e") or {}).get("value"), "frac", r.get("frac"), "GB/s", r.get("achieved"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "diff", (d.get("cpu_baseline") or {}).get("max_abs_diff_vs_gpu"))
    for k,v in (d.get("workloads") or {}).items(): print("   ", k, v.get("value"), v.get("e2e"), (v.get("roofline") or {}).get("frac"), v.get("max_abs_diff_vs_cpu"), v.get("error"))
except Exception as e:
    print("parse failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
