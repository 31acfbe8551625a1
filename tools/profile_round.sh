#!/bin/bash
# Round evidence for profiles/: plain bench line, ncu launch list and one full capture of the dominant kernel (C2).
# Run on the GPU box:  gpurun -- 'bash tools/profile_round.sh r1b'
tag=${1:-rX}
mkdir -p gpurun_out
python bench.py --steps 1000 --warmup 20 > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c2.csv \
    python bench.py --no-cpu-baseline --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flock_env -s 12 -c 1 -o gpurun_out/${tag}_c2 -f \
    python bench.py --no-cpu-baseline --steps 20 --warmup 5 --e2e-steps 2 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -c 600 gpurun_out/${tag}_bench_c2.json
