#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_fast.py -x -q -m gpu 2>&1 | tail -5
for k in 1 2; do
python bench.py --config C2k --no-cpu-baseline --no-extras --e2e-steps 2 --steps 200 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C2k', round(d['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3))"
done
python bench.py --config C2k --no-cpu-baseline --no-extras --e2e-steps 2 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C2k K=20', round(d['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3))"
