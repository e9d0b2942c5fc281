#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/fe_probe2.py > gpurun_out/s8_fe_probe.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s8_fe_launches.csv python scripts/ncu_target.py 20000 faster-evgen,no-photon-sorting > gpurun_out/s8_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_walk_kernel -c 1 -o gpurun_out/s8_fe_walk python scripts/ncu_target.py 20000 faster-evgen,no-photon-sorting > gpurun_out/s8_ncu2.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/s8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s8_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s8_bench_n1.json 2> gpurun_out/s8_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --features faster-evgen,no-photon-sorting --events 2e9 > gpurun_out/s8_bench_fe.json 2>> gpurun_out/s8_bench.err
cat gpurun_out/s8_fe_probe.txt; grep fe_ gpurun_out/s8_fe_launches.csv | tail -3 | cut -c60-250; tail -6 gpurun_out/s8_pytest.log; tail -3 gpurun_out/s8_bench.err
