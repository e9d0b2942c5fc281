#!/bin/bash
# session 49: cost of the in-kernel fold with the shipped schedule
mkdir -p gpurun_out
timeout 600 python scripts/fold_cost_probe.py > gpurun_out/s49_fold_cost.txt 2>&1
timeout 300 python scripts/fold_cost_probe.py standard-random,f32 >> gpurun_out/s49_fold_cost.txt 2>&1
cat gpurun_out/s49_fold_cost.txt
