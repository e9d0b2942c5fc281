#!/bin/bash
# session 27: compute-sanitizer on the round-2 kernels
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck python scripts/sanitize_target.py > gpurun_out/s27_memcheck.log 2>&1
timeout 1500 compute-sanitizer --tool racecheck python scripts/sanitize_target.py > gpurun_out/s27_racecheck.log 2>&1
timeout 900 compute-sanitizer --tool synccheck python scripts/sanitize_target.py > gpurun_out/s27_synccheck.log 2>&1
for t in memcheck racecheck synccheck; do echo "== $t"; grep -v "^=========$" gpurun_out/s27_$t.log | tail -32 | cut -c1-220; done
