#!/bin/bash
# session 62: unit size with the new tail
mkdir -p gpurun_out
(
timeout 300 python scripts/tail_probe.py "" --short
timeout 300 python scripts/tail_probe.py "" unit_batches=16 --short
timeout 300 python scripts/tail_probe.py "" unit_batches=12 --short
timeout 300 python scripts/tail_probe.py "" unit_batches=4 --short
) > gpurun_out/s62_tail.txt 2>&1
cat gpurun_out/s62_tail.txt
