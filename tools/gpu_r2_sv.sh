#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2_sv.log 2>&1; tail -4 gpurun_out/pytest_r2_sv.log
{
python tools/ab_bench.py
BWQ_FLAGS=32 python tools/ab_bench.py
BWQ_FLAGS=0x25000 python tools/ab_bench.py
BWQ_FLAGS=0x4a000 python tools/ab_bench.py
} > gpurun_out/ab_persist2.log 2>&1
cat gpurun_out/ab_persist2.log
timeout 600 python bench.py --workload tfim30_sv --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2_sv30.json 2> gpurun_out/bench_r2_sv30.err
BWQ_SVX_NO_FUSE=1 BWQ_SVX_NO_STRUCT=1 timeout 600 python bench.py --workload tfim30_sv --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2_sv30_plain.json 2>> gpurun_out/bench_r2_sv30.err
BWQ_SVX_NO_FUSE=1 timeout 600 python bench.py --workload tfim30_sv --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2_sv30_nofuse.json 2>> gpurun_out/bench_r2_sv30.err
python - <<'PY'
import json
for f in ("bench_r2_sv30","bench_r2_sv30_plain","bench_r2_sv30_nofuse"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("sweep_share_of_step"))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_r2.py -m gpu -x -q -k "tma or reset or default_noise" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_r2.py -m gpu -x -q -k "tma_kernel" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/sanitizer_racecheck.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sv_sweep -s 10 -c 2 -f -o gpurun_out/sv_sweep_tfim26_r2 \
   python tools/sv_bench.py 26 > gpurun_out/ncu_sv_r2.log 2>&1; tail -2 gpurun_out/ncu_sv_r2.log
