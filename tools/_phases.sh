L=${1:-build/lib_a_v9.so}
for T in 0 512 1024 2048 1536 3584 3072; do
FLOCK_LIB=$PWD/$L FLOCK_TUNE=$T python bench.py --no-cpu-baseline --steps 1000 --warmup 20 --e2e-steps 2 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$L tune=$T (bits: 512 no-expand, 1024 no-refine, 2048 no-filter)', round(d['ms_per_step']*1e3,2),'us')"
done
