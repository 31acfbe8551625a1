L=build/lib_f_pipe.so
for T in 0 32768 65536 98304 16384; do
FLOCK_LIB=$PWD/$L FLOCK_TUNE=$T python bench.py --no-cpu-baseline --steps 1000 --warmup 20 --e2e-steps 2 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$L tune=$T (32768 all compute-first, 65536 immediate, 16384 no pipe)', round(d['ms_per_step']*1e3,2),'us')"
done
FLOCK_LIB=$PWD/$L ncu --set full --clock-control none --import-source on -k regex:flock_env -s 12 -c 1 -o gpurun_out/prof_pipe -f python bench.py --no-cpu-baseline --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/ncu_pipe.log 2>&1; tail -1 gpurun_out/ncu_pipe.log
