python bench.py --no-cpu-baseline --steps 1000 --warmup 20 --e2e-steps 5 > gpurun_out/b_c2.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_c2.log').read().strip().split('\n')[-1])
print('C2 us/step', d['ms_per_step']*1e3, 'frac', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'])
PY
ncu --set full --clock-control none --import-source on -k regex:flock_env -s 12 -c 1 -o gpurun_out/prof_v11 -f python bench.py --no-cpu-baseline --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/ncu11.log 2>&1; tail -1 gpurun_out/ncu11.log
