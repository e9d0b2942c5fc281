#!/bin/bash
# session 55: unit traces of the fused kernel at the 8-GPU share, several schedules
mkdir -p gpurun_out
(
timeout 200 python scripts/unit_trace.py 125000 taper_units=-1
timeout 200 python scripts/unit_trace.py 123136 taper_units=-1
timeout 200 python scripts/unit_trace.py 125000
timeout 200 python scripts/unit_trace.py 125000 taper_units=1184 tail_singles=4736
timeout 200 python scripts/unit_trace.py 500000 taper_units=-1
) > gpurun_out/s55_trace.txt 2>&1
cat gpurun_out/s55_trace.txt
