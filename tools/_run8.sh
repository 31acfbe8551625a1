python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for L in build/lib_g_pred.so build/lib_i_swarm.so; do
for ADJ in dense none; do
FLOCK_LIB=$PWD/$L python bench.py --config C5 --steps 20 --warmup 3 --adjacency $ADJ 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$L C5 $ADJ', round(d['ms_per_step'],3),'ms', round(d['roofline']['frac'],3))"
done; done
python bench.py --no-cpu-baseline --config C3 --steps 1000 --warmup 20 --e2e-steps 2 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C3', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
