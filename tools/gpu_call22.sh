#!/bin/bash
# one --set full capture of the Flocking-v0 group kernel in steady state (a 200-step launch)
B="python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --warmup 5"
$B --config C2k --steps 200 --warmup-ms 0 > gpurun_out/x_plain_c2k200.log 2>&1 || { tail -5 gpurun_out/x_plain_c2k200.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:flock_group -s 2 -c 1 -o gpurun_out/x_c2k200 -f $B --config C2k --steps 200 --warmup-ms 0 > gpurun_out/x_ncu_c2k200.log 2>&1
ls -la gpurun_out/x_c2k200.ncu-rep
