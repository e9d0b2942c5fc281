#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/sched_probe.py 1000000 "" > gpurun_out/s3_sched_f64.txt 2>&1
timeout 600 python scripts/sched_probe.py 125000 "" > gpurun_out/s3_sched_f64_125k.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "faster_evgen or device_merge or stream_continues" -p no:cacheprovider > gpurun_out/s3_pytest_fe.log 2>&1
timeout 600 python scripts/fe_probe2.py > gpurun_out/s3_fe_probe.txt 2>&1
grep -v "^Exception\|^Traceback\|File \"\|TypeError\|^$" gpurun_out/s3_sched_f64.txt gpurun_out/s3_sched_f64_125k.txt | cut -c1-200
tail -25 gpurun_out/s3_pytest_fe.log; cat gpurun_out/s3_fe_probe.txt
