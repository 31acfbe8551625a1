#!/bin/bash
run() { python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --warmup 5 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$*', round(d['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3),'deg',d['config']['mean_degree_at_end'])"; }
for c in C2 C2k C3; do
for k in 20 200 1000; do
run --config $c --steps $k
run --config $c --steps $k --keep-warm-state
done; done
run --config C4 --steps 20
run --config C4 --steps 20 --keep-warm-state
