#!/bin/bash
# session 22: re-rolls of the walk served from four pairs tested together
mkdir -p gpurun_out
timeout 600 python scripts/fe_probe2.py > gpurun_out/s22_fe_probe.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "faster_evgen or histogram" -p no:cacheprovider > gpurun_out/s22_pytest.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s22_fe_launches.csv python scripts/ncu_target.py 20000 faster-evgen,no-photon-sorting > gpurun_out/s22_ncu1.log 2>&1
cat gpurun_out/s22_fe_probe.txt; tail -3 gpurun_out/s22_pytest.log; grep fe_ gpurun_out/s22_fe_launches.csv | tail -3 | cut -c60-250
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_walk_kernel -c 1 -o gpurun_out/s22_fe_walk python scripts/ncu_target.py 20000 faster-evgen,no-photon-sorting > gpurun_out/s22_ncu2.log 2>&1
