#!/bin/bash
# session 25: batch parts (2, 5, 10) after the fix; whole GPU suite; default bench line as a regression check of the fused kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/s25_pytest.log 2>&1
timeout 300 python scripts/default_run_timing.py > gpurun_out/s25_default_run.txt 2>&1
timeout 300 python scripts/default_run_timing.py f32 > gpurun_out/s25_default_run_f32.txt 2>&1
timeout 900 python bench.py > gpurun_out/s25_bench_n1.json 2> gpurun_out/s25_bench_n1.err
timeout 600 python bench.py --features faster-evgen,no-photon-sorting --events 2e9 > gpurun_out/s25_bench_fe_n1.json 2> gpurun_out/s25_bench_fe_n1.err
tail -5 gpurun_out/s25_pytest.log; cat gpurun_out/s25_default_run.txt gpurun_out/s25_default_run_f32.txt; cut -c1-600 gpurun_out/s25_bench_n1.json gpurun_out/s25_bench_fe_n1.json
