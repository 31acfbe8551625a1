for LN in 8 32; do for ADJ in none dense; do
python bench.py --config C5 --steps 20 --warmup 3 --adjacency $ADJ --lanes $LN 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('C5 lanes=$LN $ADJ', round(d['ms_per_step'],3),'ms', round(d['roofline']['frac'],3))"
done; done
