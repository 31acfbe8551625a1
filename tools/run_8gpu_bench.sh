#!/bin/bash
# the default bench line on G ranks only (headline C2 + extra.configs): gpurun --gpus 8 -- 'bash tools/run_8gpu_bench.sh r2g 8'
tag=${1:-rX}; G=${2:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_${G}gpu.json 2> gpurun_out/${tag}_${G}gpu.err
tail -c 400 gpurun_out/${tag}_bench_${G}gpu.json; tail -3 gpurun_out/${tag}_${G}gpu.err
