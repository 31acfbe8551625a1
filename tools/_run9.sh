python -m pytest tests -m gpu -x -q 2>&1 | tail -4
L=build/lib_j_direct.so
for T in 0 4096; do
FLOCK_LIB=$PWD/$L FLOCK_TUNE=$T python bench.py --no-cpu-baseline --steps 2000 --warmup 50 --e2e-steps 2 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C2 tune=$T (4096: out_mode 0 = direct fill + scatter)', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
for C in C4 C4x C3; do
FLOCK_LIB=$PWD/$L FLOCK_TUNE=$T python bench.py --no-cpu-baseline --config $C --steps 200 --warmup 5 --e2e-steps 2 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$C tune=$T', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done; done
