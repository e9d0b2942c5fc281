#!/bin/bash
# session 56: the shipped taper (half a wave of half units, half a wave of quarter units, two waves of single batches): tests, trace, scan
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "stream_continues or streamed_per_batch or batches_and_merged or batch_range" > gpurun_out/s56_pytest.log 2>&1
tail -2 gpurun_out/s56_pytest.log
(
timeout 200 python scripts/unit_trace.py 125000
timeout 200 python scripts/unit_trace.py 1000000
) > gpurun_out/s56_trace.txt 2>&1
(
timeout 300 python scripts/tail_probe.py ""
timeout 300 python scripts/tail_probe.py "" taper_units=-1
) > gpurun_out/s56_tail.txt 2>&1
grep -v "big units" gpurun_out/s56_trace.txt | cut -c1-400
