#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 900 -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -8 gpurun_out/r2g_pytest.log
timeout 300 python tools/shard_time.py 3,8 32 0 2>&1 | tail -2
timeout 300 python tools/shard_time.py 3,8 32 1 2>&1 | tail -2
timeout 300 python bench.py --config C2k --steps 200 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 3 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C2k K=200', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
B="python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --warmup 5"
$B --config C2k --steps 20 --warmup-ms 100 > gpurun_out/r2g_plain_c2k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flock_group -s 8 -c 1 -o gpurun_out/r2g_c2k -f \
    $B --config C2k --steps 20 --warmup-ms 100 > gpurun_out/r2g_ncu_c2k.log 2>&1
ls -la gpurun_out/r2g_c2k.ncu-rep
