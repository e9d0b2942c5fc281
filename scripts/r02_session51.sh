#!/bin/bash
# session 51: launch-size scan of the fused kernel around the 8-GPU share
mkdir -p gpurun_out
timeout 600 python scripts/tail_probe.py > gpurun_out/s51_tail.txt 2>&1
cat gpurun_out/s51_tail.txt
