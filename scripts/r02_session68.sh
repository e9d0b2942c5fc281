#!/bin/bash
# session 68: compute-sanitizer on the launch shapes added in sessions 52-64
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> gpurun_out/s68_sanitizer.txt
  timeout 600 compute-sanitizer --tool $tool python scripts/sanitize_target.py --schedule >> gpurun_out/s68_sanitizer.txt 2>&1
done
grep -v "^$" gpurun_out/s68_sanitizer.txt | cut -c1-220
