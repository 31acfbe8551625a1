#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_swarm.py tests/test_gpu_group.py -x -q -m gpu 2>&1 | tail -5
python tools/shard_time.py 3,8 32 0 0x10000000 0 2>&1 | grep "swarm step" | awk 'NR%15==1' | head -12
python tools/shard_time.py 3,8 32 0 0 0 2>&1 | tail -6
