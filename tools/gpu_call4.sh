#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 900 -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
for C in C2 C2k C3; do for K in 20 200; do
timeout 300 python bench.py --config $C --steps $K --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 3 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$C K=$K', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done; done
timeout 300 python tools/shard_time.py 3,8 32 > gpurun_out/r2d_shard_plain.log 2>&1 && cat gpurun_out/r2d_shard_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:flock_swarm_kernel -s 8 -c 1 -o gpurun_out/r2d_shard_dense -f python tools/shard_time.py 3,8 32 > gpurun_out/r2d_ncu_shard.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flock_swarm_kernel -s 120 -c 1 -o gpurun_out/r2d_shard_none -f python tools/shard_time.py 3,8 32 > gpurun_out/r2d_ncu_shard2.log 2>&1
ls -la gpurun_out/r2d_*.ncu-rep
