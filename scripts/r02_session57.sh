#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/unit_trace.py 125000 > gpurun_out/s57_trace.txt 2>&1
grep -v "big units" gpurun_out/s57_trace.txt | cut -c1-300
