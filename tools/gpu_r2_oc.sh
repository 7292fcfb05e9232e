#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_onchip_gpu.py -m gpu -x -q 2>&1 | tail -30
timeout 300 python tools/cfg1_breakdown.py 2>&1 | grep -v host_threads | tee gpurun_out/cfg1_breakdown_oc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_onchip_kernel -s 19 -c 1 -o gpurun_out/onchip_cfg1 -f python tools/cfg1_breakdown.py > gpurun_out/ncu_onchip.log 2>&1; tail -2 gpurun_out/ncu_onchip.log
