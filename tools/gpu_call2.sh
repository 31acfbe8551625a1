#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 900 -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 300 python tools/shard_time.py 3,8 32 2>&1 | tail -4
timeout 300 python tools/shard_time.py 3,8 16 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_short.json 2> gpurun_out/r2b_bench_short.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_short.json').read().strip().split('\n')[-1])
print('short: ms/step', round(d['ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3), 'per-step api', d['extra']['per_step_api'])
for k,v in d['extra']['configs'].items(): print('  ',k,{kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms_per_step','frac','error','adjacency_none_ms_per_step')})
PY
for K in 20 200 1000; do
timeout 300 python bench.py --steps $K --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 3 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('K=$K', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done
for G in 1 2 3 5; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 3 --tune $((G<<16)) 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('K=20 G=$G', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
timeout 300 python bench.py --steps 1000 --warmup 5 --no-extras --no-cpu-baseline --e2e-steps 3 --tune $(( (G<<16) | 0x4000000 )) 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('per-step launches G=$G', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2b_memcheck.log python tools/sanitize_cases.py > gpurun_out/r2b_memcheck.out 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r2b_memcheck.out; tail -8 gpurun_out/r2b_memcheck.log
