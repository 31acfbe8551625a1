#!/bin/bash
# first GPU call of round 2: parity suite, store micro-benchmark, short + long bench lines, A/B vs the round-1 kernel path
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 900 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -40 gpurun_out/r2a_pytest.log
timeout 120 ./tools/ubench_store.bin > gpurun_out/r2a_store.txt 2>&1; cat gpurun_out/r2a_store.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_short.json 2> gpurun_out/r2a_bench_short.err; tail -c 3000 gpurun_out/r2a_bench_short.json; tail -5 gpurun_out/r2a_bench_short.err
timeout 300 python bench.py --steps 1000 --warmup 20 --no-extras --no-cpu-baseline --e2e-steps 5 > gpurun_out/r2a_bench_long.json 2> gpurun_out/r2a_bench_long.err; tail -c 1500 gpurun_out/r2a_bench_long.json
# round-1 path (flock_env_kernel, K launches): tune bit 24
timeout 300 python bench.py --steps 1000 --warmup 20 --no-extras --no-cpu-baseline --e2e-steps 5 --tune 0x1000000 > gpurun_out/r2a_bench_long_envkernel.json 2>&1; tail -c 700 gpurun_out/r2a_bench_long_envkernel.json
# per-step launches of the group kernel, with and without PDL
timeout 300 python bench.py --steps 1000 --warmup 20 --no-extras --no-cpu-baseline --e2e-steps 5 --tune 0x4000000 > gpurun_out/r2a_bench_long_nofuse.json 2>&1; tail -c 700 gpurun_out/r2a_bench_long_nofuse.json
timeout 300 python bench.py --steps 1000 --warmup 20 --no-extras --no-cpu-baseline --e2e-steps 5 --tune 0x6000000 > gpurun_out/r2a_bench_long_nofuse_nopdl.json 2>&1; tail -c 700 gpurun_out/r2a_bench_long_nofuse_nopdl.json
for G in 1 2 3 4; do
timeout 300 python bench.py --steps 1000 --warmup 20 --no-extras --no-cpu-baseline --e2e-steps 5 --tune $((G<<16)) 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('G=$G', round(d['ms_per_step']*1e3,2),'us', round(d['roofline']['frac'],3))"
done
