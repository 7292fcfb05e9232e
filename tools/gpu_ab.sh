#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/quick_bench.py 2>&1 | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'frac',round(d['roofline']['frac'],3))"
