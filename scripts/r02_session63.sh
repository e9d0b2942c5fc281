#!/bin/bash
# session 63: units of 16 batches against units of 8, full launch-size scans on the same box
mkdir -p gpurun_out
(
timeout 300 python scripts/tail_probe.py "" unit_batches=16
timeout 300 python scripts/tail_probe.py ""
) > gpurun_out/s63_tail.txt 2>&1
