#!/bin/bash
run() { python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --warmup 5 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$*', round(d['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3),'deg',d['config']['mean_degree_at_end'])"; }
for g in 1 2 3 4 5; do
run --config C2 --steps 200 --tune $((g<<16))
run --config C2k --steps 200 --tune $((g<<16))
done
for g in 2 4 5 6 8 10; do
run --config C3 --steps 200 --tune $((g<<16))
done
