#!/bin/bash
# session 45: faster-evgen segment length sweep
mkdir -p gpurun_out
timeout 900 python scripts/fe_seg_probe.py 4e9 > gpurun_out/s45_fe_seg.txt 2>&1
cat gpurun_out/s45_fe_seg.txt; nvidia-smi --query-gpu=memory.used,memory.total --format=csv
