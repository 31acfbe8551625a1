#!/bin/bash
# phase costs of one config through the kernels' diagnostic bits (FlockParams.reserved: bits 8.. skip fill / scatter / refine,
# bit 29 exact-order fill as plain stores instead of TMA bulk stores): timing only, the skipping variants compute wrong results.
# usage: CFG=C4x K=10 VARIANTS="0 0x100 0x300 0x700 0x20000000" bash tools/diag_phases.sh
CFG=${CFG:-C4x}; K=${K:-10}
run() { python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --warmup 5 --config $CFG --steps $K "$@" 2>gpurun_out/diag_err.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$CFG $*', round(d['ms_per_step']*1e3,2),'us frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print('FAILED', '$*'); print(open('gpurun_out/diag_err.log').read()[-1500:])"; }
for v in ${VARIANTS:-0 0x100 0x300 0x700}; do run --tune $v; done
[ -n "$NONE" ] && run --adjacency none
