#!/bin/bash
# Round evidence for profiles/ (one gpurun call; every ncu run follows a plain run of the same command line):
#   plain bench line, ncu launch list of the headline command, one --set full capture per dominant kernel (C2, C2k, C4, C5).
# gpurun -- 'bash tools/profile_round.sh r2'
tag=${1:-rX}
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --warmup 5"
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_driver_cmd.json 2> gpurun_out/${tag}_bench_driver_cmd.err || exit 1
python bench.py --steps 1000 --warmup 20 --no-extras > gpurun_out/${tag}_bench_c2_long.json 2>> gpurun_out/${tag}_bench_driver_cmd.err || exit 1
$B --steps 20 > gpurun_out/${tag}_plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c2.csv \
    $B --steps 20 > gpurun_out/${tag}_ncu_launches.log 2>&1
$B --steps 20 --warmup-ms 0 > gpurun_out/${tag}_plain_c2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flock_group -s 2 -c 1 -o gpurun_out/${tag}_c2 -f \
    $B --steps 20 --warmup-ms 0 > gpurun_out/${tag}_ncu_c2.log 2>&1
$B --config C2k --steps 20 --warmup-ms 0 > gpurun_out/${tag}_plain_c2k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flock_group -s 2 -c 1 -o gpurun_out/${tag}_c2k -f \
    $B --config C2k --steps 20 --warmup-ms 0 > gpurun_out/${tag}_ncu_c2k.log 2>&1
$B --config C4 --steps 3 --warmup-ms 0 > gpurun_out/${tag}_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flock_env_kernel -s 4 -c 1 -o gpurun_out/${tag}_c4 -f \
    $B --config C4 --steps 3 --warmup-ms 0 > gpurun_out/${tag}_ncu_c4.log 2>&1
$B --config C5 --steps 3 > gpurun_out/${tag}_plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:flock_swarm_kernel -s 4 -c 1 -o gpurun_out/${tag}_c5 -f \
    $B --config C5 --steps 3 > gpurun_out/${tag}_ncu_c5.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
tail -c 400 gpurun_out/${tag}_bench_c2_long.json
