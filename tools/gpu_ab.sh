#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/quick_bench.py 2>&1 | grep -E "brick10 6 2 0|tfim 1[23] 6 2" | cut -c1-200
for w in brick10_guadalupe_twirl tfim12_dm; do
timeout 300 python bench.py --no-cpu-baseline --workload $w --steps 3 --warmup 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3))"
done
