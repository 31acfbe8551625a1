#!/bin/bash
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
bash tools/profile_round.sh r2i c4x
