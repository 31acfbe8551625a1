#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_fast.py tests/test_gpu_swarm.py -x -q -m gpu 2>&1 | tail -4
run() { python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --warmup 5 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$*', round(d['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3),'deg',d['config']['mean_degree_at_end'])"; }
run --config C2 --steps 200
run --config C2 --steps 20
run --config C2k --steps 200
run --config C3 --steps 200
run --config C4 --steps 20
