#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/shard_time.py 3,8 32 0 > gpurun_out/r2h_shard_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2h_shard_launches.csv python tools/shard_time.py 3,8 32 0 > gpurun_out/r2h_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2h_shard_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:40]: print(r[ki][:60], r[vi])
PY
