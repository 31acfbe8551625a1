#!/bin/bash
# Round evidence for profiles/.  One ncu run per gpurun call, each after a plain run of the same command line exited 0:
#   gpurun -- 'bash tools/profile_round.sh r2 bench'      plain bench lines (driver command, long run)
#   gpurun -- 'bash tools/profile_round.sh r2 launches'   ncu launch list of the headline command
#   gpurun -- 'bash tools/profile_round.sh r2 c2|c2k|c3|c4|c4x|c5'   one --set full capture of that config's dominant kernel
tag=${1:-rX}; stage=${2:-bench}
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --warmup 5"
full() {   # full <name> <kernel regex> <skip> <bench args...>
    name=$1; kre=$2; skip=$3; shift 3
    $B "$@" > gpurun_out/${tag}_plain_${name}.log 2>&1 || { tail -5 gpurun_out/${tag}_plain_${name}.log; exit 1; }
    ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -o gpurun_out/${tag}_${name} -f \
        $B "$@" > gpurun_out/${tag}_ncu_${name}.log 2>&1
    ls -la gpurun_out/${tag}_${name}.ncu-rep
}
case $stage in
bench)
    python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_driver_cmd.json 2> gpurun_out/${tag}_bench_driver_cmd.err || exit 1
    python bench.py --steps 1000 --warmup 20 --no-extras > gpurun_out/${tag}_bench_c2_long.json 2>> gpurun_out/${tag}_bench_driver_cmd.err || exit 1
    python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench_driver_cmd.err
    tail -c 600 gpurun_out/${tag}_bench_driver_cmd.json; tail -c 300 gpurun_out/${tag}_bench_c2_long.json ;;
launches)
    $B --steps 20 > gpurun_out/${tag}_plain_c2.log 2>&1 || exit 1
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c2.csv \
        $B --steps 20 > gpurun_out/${tag}_ncu_launches.log 2>&1
    tail -5 gpurun_out/${tag}_launches_c2.csv ;;
c2)  full c2 flock_group 2 --steps 20 --warmup-ms 0 ;;
c2k) full c2k flock_group 2 --config C2k --steps 20 --warmup-ms 0 ;;
c3)  full c3 flock_group 2 --config C3 --steps 20 --warmup-ms 0 ;;
c4)  full c4 flock_env_kernel 4 --config C4 --steps 3 --warmup-ms 0 ;;
c4x) full c4x flock_env_kernel 4 --config C4x --steps 3 --warmup-ms 0 ;;
c5)  full c5 flock_swarm_kernel 4 --config C5 --steps 3 ;;
*) echo "unknown stage $stage"; exit 2 ;;
esac
