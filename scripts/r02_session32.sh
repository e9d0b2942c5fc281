#!/bin/bash
# session 32: f32 sin cos quadrant split on the float adder (packed for the two events of a lane)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/s32_pytest.log 2>&1
timeout 300 python scripts/f32_golden_report.py > gpurun_out/s32_f32_golden_digits.txt 2>&1
timeout 600 python bench.py --features standard-random,f32 --no-cpu-baseline > gpurun_out/s32_bench_f32xo.json 2> gpurun_out/s32_bench_f32xo.err
timeout 600 python bench.py --features f32 --no-cpu-baseline > gpurun_out/s32_bench_f32.json 2> gpurun_out/s32_bench_f32.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simulate_kernel_x2 -c 1 -o gpurun_out/s32_f32x2_kernel python scripts/ncu_target.py 200000 standard-random,f32 0 1 > gpurun_out/s32_ncu.log 2>&1
tail -3 gpurun_out/s32_pytest.log; for f in f32xo f32; do python -c "
import json,sys; d=json.loads(open('gpurun_out/s32_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"; done; cmp gpurun_out/s32_f32_golden_digits.txt profiles/r02_f32_golden_digits.txt && echo "f32 golden report unchanged"
