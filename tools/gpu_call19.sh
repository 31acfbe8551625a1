#!/bin/bash
python tools/soak_group.py 120 7 > gpurun_out/r2_soak_group.txt 2>&1; echo "soak rc=$?"; tail -5 gpurun_out/r2_soak_group.txt; grep -c OK gpurun_out/r2_soak_group.txt; grep DIFF gpurun_out/r2_soak_group.txt | head
