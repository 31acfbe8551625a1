#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 900 -p no:cacheprovider > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -8 gpurun_out/r2e_pytest.log
for C in C2 C2k C3; do for K in 20 200; do
timeout 300 python bench.py --config $C --steps $K --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 3 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$C K=$K', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done; done
for G in 1 2 3 4 5; do
timeout 300 python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 3 --tune $((G<<16)) 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C2 K=200 G=$G', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done
timeout 300 python bench.py --steps 1000 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 3 --tune 0x4000000 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C2 per-step launches', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
