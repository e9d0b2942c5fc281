#!/bin/bash
# session 53: full launch-size scans of a few taper settings
mkdir -p gpurun_out
(
timeout 300 python scripts/tail_probe.py "" taper_units=1184
timeout 300 python scripts/tail_probe.py "" taper_units=1184 tail_singles=2368
timeout 300 python scripts/tail_probe.py "" taper_units=592
timeout 300 python scripts/tail_probe.py "" taper_units=1184 tail_singles=3552
timeout 300 python scripts/tail_probe.py "" taper_units=1184 ramp_units=2368
timeout 300 python scripts/tail_probe.py "" taper_units=1776
) > gpurun_out/s53_tail.txt 2>&1
grep -c "n =" gpurun_out/s53_tail.txt
