#!/bin/bash
# split step barrier: correctness of the group kernel, then A/B against the previous build
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_fast.py -x -q -m gpu > gpurun_out/t20_pytest.log 2>&1; tail -4 gpurun_out/t20_pytest.log
run() { python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --warmup 5 "$@" 2>gpurun_out/t20_err.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$FLOCK_LIB $*', round(d['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3),'deg',d['config']['mean_degree_at_end'])
except Exception as e: print('FAILED', '$FLOCK_LIB $*')"; tail -3 gpurun_out/t20_err.log; }
for rep in 1 2; do for L in build/lib_*.so; do
export FLOCK_LIB=$PWD/$L
run --config C2 --steps 20
run --config C2 --steps 200
run --config C2k --steps 200
run --config C3 --steps 200
done; done
