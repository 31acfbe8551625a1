#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_fast.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -5
run() { python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --warmup 5 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$*', round(d['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3),'deg',d['config']['mean_degree_at_end'])"; }
for c in C2 C2k C3; do
for k in 20 200; do
run --config $c --steps $k
run --config $c --steps $k --tune 0x20000000
done; done
