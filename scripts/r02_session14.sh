#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/fe_probe2.py > gpurun_out/s14_fe_probe.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "faster_evgen or histogram" -p no:cacheprovider > gpurun_out/s14_pytest.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s14_fe_launches.csv python scripts/ncu_target.py 100000 faster-evgen,no-photon-sorting > gpurun_out/s14_ncu1.log 2>&1
cat gpurun_out/s14_fe_probe.txt; tail -5 gpurun_out/s14_pytest.log
