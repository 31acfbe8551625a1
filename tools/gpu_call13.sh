#!/bin/bash
for t in 0 0x2000; do
python bench.py --config C4x --no-cpu-baseline --no-extras --e2e-steps 2 --steps 20 --warmup 5 --tune $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C4x tune $t', round(d['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3))"
done
