#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_onchip_gpu.py -m gpu -x -q 2>&1 | tail -30
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2_oc.log 2>&1; tail -5 gpurun_out/pytest_r2_oc.log
for l in "" tools/_build/libbwq_ocw2.so tools/_build/libbwq_ocw4.so; do
  echo "== lib $l"; BWQ_LIB=$l timeout 300 python tools/cfg1_breakdown.py 2>&1 | grep -v host_threads
done | tee gpurun_out/cfg1_breakdown_oc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_onchip_kernel -s 2 -c 1 -o gpurun_out/onchip_cfg1 -f python tools/cfg1_breakdown.py > gpurun_out/ncu_onchip.log 2>&1; tail -2 gpurun_out/ncu_onchip.log
