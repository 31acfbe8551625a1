#!/bin/bash
# 8-GPU evidence: C5 swarm rows sharded (dense / none), sharded == single-GPU check, optionally C2 / C3 env-sharded.
# gpurun --gpus 8 -- 'bash tools/run_8gpu.sh r1c 8 [all]'
tag=${1:-rX}; G=${2:-8}; what=${3:-swarm}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
err=gpurun_out/${tag}_${G}gpu.err
if [ "$what" = "all" ]; then
$TR bench.py --gpus $G --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/${tag}_c2_${G}gpu.json 2> $err
$TR bench.py --gpus $G --config C3 --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/${tag}_c3_${G}gpu.json 2>> $err
fi
$TR bench.py --gpus $G --config C5 --steps 100 --warmup 5 --adjacency dense > gpurun_out/${tag}_c5_${G}gpu_dense.json 2>> $err
$TR bench.py --gpus $G --config C5 --steps 100 --warmup 5 --adjacency none > gpurun_out/${tag}_c5_${G}gpu_none.json 2>> $err
$TR bench.py --gpus $G --config C5 --steps 100 --warmup 5 --adjacency dense --lanes 32 > gpurun_out/${tag}_c5_${G}gpu_dense_l32.json 2>> $err
$TR tools/swarm_check.py 2048 > gpurun_out/${tag}_swarm_check_${G}gpu.txt 2>> $err
for g in gpurun_out/${tag}_c*_${G}gpu*.json; do python - "$g" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1]); print(sys.argv[1].split('/')[-1], round(d['ms_per_step']*1e3,2),'us', '%.3e'%d['value'], d['roofline']['frac'])
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
grep swarm gpurun_out/${tag}_swarm_check_${G}gpu.txt; tail -3 $err
