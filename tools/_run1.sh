set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --no-cpu-baseline --steps 1000 --warmup 20 > gpurun_out/b_c2.log 2>&1; tail -c 1200 gpurun_out/b_c2.log
ncu --set full --clock-control none --import-source on -k regex:flock_env -s 12 -c 1 -o gpurun_out/prof_v10 -f python bench.py --no-cpu-baseline --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/ncu10.log 2>&1; tail -3 gpurun_out/ncu9.log
