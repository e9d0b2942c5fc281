#!/bin/bash
# session 66: ncu --set full of the bench's own launch with the final build (default f64, 1e6 batches, in-kernel fold)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -c 1 -o gpurun_out/s66_bench_kernel python scripts/ncu_target.py 1000000 "" 0 1 > gpurun_out/s66_ncu_kernel.log 2>&1
tail -2 gpurun_out/s66_ncu_kernel.log
