#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/sched_probe.py 1000000 "" > gpurun_out/s2_sched_f64.txt 2>&1
timeout 600 python scripts/sched_probe.py 125000 "" > gpurun_out/s2_sched_f64_125k.txt 2>&1
timeout 600 python scripts/sched_probe.py 1000000 "standard-random,f32" > gpurun_out/s2_sched_f32xo.txt 2>&1
cat gpurun_out/s2_sched_f64.txt gpurun_out/s2_sched_f64_125k.txt gpurun_out/s2_sched_f32xo.txt
