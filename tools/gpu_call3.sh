#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 900 -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -8 gpurun_out/r2c_pytest.log
timeout 300 python tools/shard_time.py 3,8 32 2>&1 | tail -2
for T in 0 0x8000000; do for C in C2k C2 C3; do
timeout 300 python bench.py --config $C --steps 200 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 3 --tune $T 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$C tune=$T', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done; done
bash tools/profile_round.sh r2c
