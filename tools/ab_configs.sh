#!/bin/bash
# A/B of the builds under build/ (tools/mkvariant.py) over configs and step counts: CFGS="C2 C4x" KS="20 200" REPS=2 bash tools/ab_configs.sh
run() { python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --warmup 5 "$@" 2>gpurun_out/ab_err.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$FLOCK_LIB $*', round(d['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3),'deg',d['config']['mean_degree_at_end'])
except Exception as e: print('FAILED', '$FLOCK_LIB $*')"; tail -3 gpurun_out/ab_err.log; }
for rep in $(seq 1 ${REPS:-1}); do for L in build/lib_*.so; do
export FLOCK_LIB=$PWD/$L
for cfg in ${CFGS:-C2 C2k C3}; do
for k in ${KS:-20 200}; do run --config $cfg --steps $k; done
done; done; done
