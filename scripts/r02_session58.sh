#!/bin/bash
# session 58: big units end on a band boundary (multiple of 4 x SMs units): scans with and without, with and without taper
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "stream_continues or streamed_per_batch or batches_and_merged or batch_range" > gpurun_out/s58_pytest.log 2>&1
tail -2 gpurun_out/s58_pytest.log
(
timeout 300 python scripts/tail_probe.py ""
timeout 300 python scripts/tail_probe.py "" align_units=0
timeout 300 python scripts/tail_probe.py "" taper_units=-1
timeout 300 python scripts/tail_probe.py "" taper_units=-1 align_units=0
) > gpurun_out/s58_tail.txt 2>&1
