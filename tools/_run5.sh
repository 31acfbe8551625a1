L=build/lib_f_pipe.so
for T in 16384 16388 16390 16387 16389; do
FLOCK_LIB=$PWD/$L FLOCK_TUNE=$T python bench.py --no-cpu-baseline --steps 1000 --warmup 20 --e2e-steps 2 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$L tune=$T (low byte = warps per CTA; 16384 = no pipe)', round(d['ms_per_step']*1e3,2),'us')"
done
