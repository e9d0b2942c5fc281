#!/bin/bash
# GPU session 1 of round 2: parity suite on the new schedule / fold, FP64 peak probe variants, first bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/s1_smi.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_far_streams.py -p no:cacheprovider > gpurun_out/s1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
timeout 120 scripts/micro/_build/peak64 > gpurun_out/s1_peak64.txt 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/s1_bench_n1.json 2> gpurun_out/s1_bench_n1.err
timeout 300 python scripts/f32_golden_report.py > gpurun_out/s1_f32_golden.txt 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --features standard-random,f32 --no-cpu-baseline > gpurun_out/s1_bench_f32.json 2>> gpurun_out/s1_bench_n1.err
timeout 300 python bench.py --steps 3 --warmup 3 --events 1.25e9 --no-cpu-baseline > gpurun_out/s1_bench_125k.json 2>> gpurun_out/s1_bench_n1.err
tail -3 gpurun_out/s1_pytest.log; cat gpurun_out/s1_bench_n1.json | head -c 1500
