#!/bin/bash
# session 26: default-run record after batch parts + lean context creation; ncu of the faster-evgen physics kernel after the rsqrt fusion
mkdir -p gpurun_out
timeout 300 python scripts/default_run_timing.py > gpurun_out/s26_default_run.txt 2>&1
timeout 300 python scripts/default_run_timing.py f32 > gpurun_out/s26_default_run_f32.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s26_fe_launches.csv python scripts/ncu_target.py 20000 faster-evgen,no-photon-sorting > gpurun_out/s26_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_physics_kernel -c 1 -o gpurun_out/s26_fe_phys python scripts/ncu_target.py 20000 faster-evgen,no-photon-sorting > gpurun_out/s26_ncu2.log 2>&1
cat gpurun_out/s26_default_run.txt gpurun_out/s26_default_run_f32.txt; grep fe_ gpurun_out/s26_fe_launches.csv | tail -6 | cut -c60-250
