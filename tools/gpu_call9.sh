#!/bin/bash
python tools/shard_time.py 3,8 32 0 0x10000000 0 2>&1 | grep "swarm step" | awk 'NR%15==1' | head -16
