#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/nocuts_probe.py > gpurun_out/s19_nocuts.txt 2>&1
timeout 300 python scripts/ramp_probe.py > gpurun_out/s19_ramp.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -c 1 -o gpurun_out/s19_bench_kernel python scripts/ncu_target.py 1000000 "" 0 1 > gpurun_out/s19_ncu_kernel.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s19_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s19_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:simulate_kernel_x2 -c 1 -o gpurun_out/s19_f32x2_kernel python scripts/ncu_target.py 200000 "standard-random,f32" 0 1 > gpurun_out/s19_ncu_f32.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/s19_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s19_pytest.log
cat gpurun_out/s19_nocuts.txt gpurun_out/s19_ramp.txt; tail -4 gpurun_out/s19_pytest.log; ls -la gpurun_out | grep s19
