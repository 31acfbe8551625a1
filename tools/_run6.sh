python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for L in build/lib_g_pred.so build/lib_h_bigrows.so; do
for C in C4 C4x; do
FLOCK_LIB=$PWD/$L python bench.py --no-cpu-baseline --config $C --steps 200 --warmup 5 --e2e-steps 2 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$L $C', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done; done
