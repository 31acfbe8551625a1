for rep in 1 2; do for L in build/lib_*.so; do
FLOCK_LIB=$PWD/$L python bench.py --no-cpu-baseline --steps 2000 --warmup 50 --e2e-steps 2 ${AB_ARGS} 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$L', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done; done
