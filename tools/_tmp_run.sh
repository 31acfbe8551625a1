python -m pytest tests/test_gpu_swarm.py -x -q 2>&1 | tail -5
for ADJ in dense none; do
python bench.py --config C5 --steps 20 --warmup 3 --adjacency $ADJ 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C5 $ADJ', round(d['ms_per_step'],3),'ms', round(d['roofline']['frac'],3))"
done
