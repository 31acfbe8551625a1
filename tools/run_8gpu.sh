#!/bin/bash
# 8-GPU evidence: the streaming swarm on real ranks (bit-equality at N = 8192 and 65 536, plus the env-resident N = 2048), the
# default bench line on G ranks (headline C2 + extra.configs incl. C3 = 16384/G envs per GPU and C5 with the gather delta).
# gpurun --gpus 8 -- 'bash tools/run_8gpu.sh r2 8'
tag=${1:-rX}; G=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
err=gpurun_out/${tag}_${G}gpu.err
# insurance: the parity suite must be green on this build before 8 GPUs are spent on it
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${tag}_pytest_${G}gpu.log 2>&1 || { tail -20 gpurun_out/${tag}_pytest_${G}gpu.log; exit 1; }
tail -2 gpurun_out/${tag}_pytest_${G}gpu.log
$TR tools/swarm_check.py 2048 8192 65536 > gpurun_out/${tag}_swarm_check_${G}gpu.txt 2> $err
cat gpurun_out/${tag}_swarm_check_${G}gpu.txt
$TR bench.py --gpus $G --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_${G}gpu.json 2>> $err
$TR bench.py --gpus $G --config C5 --steps 100 --warmup 5 --adjacency dense > gpurun_out/${tag}_c5_${G}gpu_dense.json 2>> $err
$TR bench.py --gpus $G --config C5 --steps 100 --warmup 5 --adjacency none > gpurun_out/${tag}_c5_${G}gpu_none.json 2>> $err
python - <<'PY' gpurun_out/${tag}_bench_${G}gpu.json gpurun_out/${tag}_c5_${G}gpu_dense.json gpurun_out/${tag}_c5_${G}gpu_none.json
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f.split('/')[-1], round(d['ms_per_step']*1e3,2),'us', '%.3e'%d['value'], round(d['roofline']['frac'],3))
        for k,v in (d.get('extra') or {}).get('configs',{}).items():
            print('   ', k, {kk: (round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms_per_step','frac','error','skip_gather_ms_per_step','adjacency_none_ms_per_step')})
    except Exception as e: print(f, 'FAILED', e)
PY
tail -3 $err
