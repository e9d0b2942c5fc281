#!/bin/bash
# session 21: physics kernel with one reciprocal square root per photon; how walk and physics share the SMs
mkdir -p gpurun_out
timeout 600 python scripts/fe_probe2.py > gpurun_out/s21_fe_probe.txt 2>&1
timeout 600 python scripts/fe_corun_probe.py > gpurun_out/s21_fe_corun.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "faster_evgen or histogram" -p no:cacheprovider > gpurun_out/s21_pytest.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s21_fe_launches.csv python scripts/ncu_target.py 20000 faster-evgen,no-photon-sorting > gpurun_out/s21_ncu1.log 2>&1
cat gpurun_out/s21_fe_probe.txt gpurun_out/s21_fe_corun.txt; tail -3 gpurun_out/s21_pytest.log; grep fe_ gpurun_out/s21_fe_launches.csv | tail -3 | cut -c60-250
