#!/bin/bash
# A/B of sweep-kernel builds under tools/_build against the default build
mkdir -p gpurun_out
for i in 1 2; do
  timeout 300 python tools/ab_bench.py
  for l in tools/_build/libbwq_*.so; do BWQ_LIB=$l timeout 300 python tools/ab_bench.py; done
done 2>&1 | tee gpurun_out/ab_r2e.log
