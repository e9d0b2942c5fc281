#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/fe_units_probe.py > gpurun_out/s69_fe_units.txt 2>&1
cat gpurun_out/s69_fe_units.txt
