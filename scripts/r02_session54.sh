#!/bin/bash
# session 54: taper + unit-count alignment of the dynamic schedule: tests, then launch-size scans
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "stream_continues or streamed_per_batch or batches_and_merged or batch_range or far_batches" > gpurun_out/s54_pytest.log 2>&1
tail -2 gpurun_out/s54_pytest.log
(
timeout 300 python scripts/tail_probe.py ""
timeout 300 python scripts/tail_probe.py "" align_units=0
timeout 300 python scripts/tail_probe.py "" taper_units=1 tail_singles=9472
timeout 300 python scripts/tail_probe.py "" taper_units=2368 tail_singles=4736
timeout 300 python scripts/tail_probe.py "" taper_units=1184 tail_singles=4736
) > gpurun_out/s54_tail.txt 2>&1
grep -c "n =" gpurun_out/s54_tail.txt
