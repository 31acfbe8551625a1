for LN in 0 1 2 4; do
python bench.py --no-cpu-baseline --config C3 --steps 1000 --warmup 20 --e2e-steps 2 --env-lanes $LN 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C3 lanes=$LN', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done
for T in 4 6; do
FLOCK_TUNE=$T python bench.py --no-cpu-baseline --config C3 --steps 1000 --warmup 20 --e2e-steps 2 --env-lanes 1 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C3 lanes=1 warps/CTA=$T', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done
