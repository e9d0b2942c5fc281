#!/bin/bash
# session 64: big-unit size chosen by launch length (1 .. 16) against fixed units of 8, short to long launches
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "stream_continues or streamed_per_batch or batches_and_merged or batch_range or launch_shape or default_run_matches" > gpurun_out/s64_pytest.log 2>&1
tail -2 gpurun_out/s64_pytest.log
S="n=1000,3000,6000,12000,20000,30000,40000,60000,90000,125000,250000,500000"
(
timeout 300 python scripts/tail_probe.py "" $S
timeout 300 python scripts/tail_probe.py "" unit_batches=8 $S
timeout 300 python scripts/tail_probe.py "f32" $S
timeout 300 python scripts/tail_probe.py "f32" unit_batches=8 $S
) > gpurun_out/s64_tail.txt 2>&1
cat gpurun_out/s64_tail.txt
