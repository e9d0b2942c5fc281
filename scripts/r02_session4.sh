#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/s4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s4_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4_fe_launches.csv python scripts/ncu_target.py 20000 faster-evgen,no-photon-sorting > gpurun_out/s4_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_walk_kernel -c 1 -o gpurun_out/s4_fe_walk python scripts/ncu_target.py 20000 faster-evgen,no-photon-sorting > gpurun_out/s4_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fe_physics_kernel -c 1 -o gpurun_out/s4_fe_phys python scripts/ncu_target.py 20000 faster-evgen,no-photon-sorting > gpurun_out/s4_ncu3.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/s4_bench_n1.json 2> gpurun_out/s4_bench_n1.err
timeout 300 python bench.py --steps 5 --warmup 3 --features standard-random,f32 --no-cpu-baseline > gpurun_out/s4_bench_f32.json 2>> gpurun_out/s4_bench_n1.err
timeout 300 python bench.py --steps 3 --warmup 3 --events 1.25e9 --no-cpu-baseline > gpurun_out/s4_bench_125k.json 2>> gpurun_out/s4_bench_n1.err
timeout 300 python scripts/default_run_timing.py > gpurun_out/s4_default_run.txt 2>&1
tail -30 gpurun_out/s4_pytest.log; cat gpurun_out/s4_default_run.txt; ls -la gpurun_out/
