#!/bin/bash
# correctness of the group kernel (tests + randomised cross-check against the env kernel), then A/B of the builds under build/
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_fast.py tests/test_gpu_api.py -x -q -m gpu > gpurun_out/t21_pytest.log 2>&1; tail -4 gpurun_out/t21_pytest.log
python tools/soak_group.py ${SOAK:-40} 11 > gpurun_out/t21_soak.txt 2>&1; echo "soak rc=$?"; tail -2 gpurun_out/t21_soak.txt; grep DIFF gpurun_out/t21_soak.txt | head -5
run() { python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --warmup 5 "$@" 2>gpurun_out/t21_err.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$FLOCK_LIB $*', round(d['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3),'deg',d['config']['mean_degree_at_end'])
except Exception as e: print('FAILED', '$FLOCK_LIB $*')"; tail -3 gpurun_out/t21_err.log; }
for rep in 1 2; do for L in build/lib_*.so; do
export FLOCK_LIB=$PWD/$L
for cfg in ${CFGS:-C2 C2k C3}; do
run --config $cfg --steps 20
run --config $cfg --steps 200
done; done; done
