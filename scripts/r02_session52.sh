#!/bin/bash
# session 52: taper stages (half units, quarter units) between the big units and the single batches of the dynamic schedule
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "stream_continues or streamed_per_batch or batches_and_merged or batch_range" > gpurun_out/s52_pytest.log 2>&1
tail -2 gpurun_out/s52_pytest.log
(
timeout 300 python scripts/tail_probe.py "" --short
timeout 300 python scripts/tail_probe.py "" taper_units=-1 --short
timeout 300 python scripts/tail_probe.py "" tail_singles=2368 --short
timeout 300 python scripts/tail_probe.py "" tail_singles=7104 --short
timeout 300 python scripts/tail_probe.py "" taper_units=1184 --short
timeout 300 python scripts/tail_probe.py "" taper_units=4736 --short
timeout 300 python scripts/tail_probe.py "" taper_units=-1 tail_singles=18944 --short
) > gpurun_out/s52_tail.txt 2>&1
cat gpurun_out/s52_tail.txt
