#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lib in "" /root/repo/ml_qem_b200/lib/libbwq_b4.so; do
  echo "== BWQ_LIB=$lib"
  BWQ_LIB=$lib timeout 300 python tools/quick_bench.py 2>&1 | grep -E "brick10 6 2 0|tfim 1[23] 6 2" | cut -c1-150
done
