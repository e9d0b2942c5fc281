#!/bin/bash
# session 47: f32 sin / cos as ONE sin.approx / cos.approx on the reflected angle (|x| <= pi) -- every f32 bound re-measured,
# and the bench lines next to the quadrant form (libtp3_quadrant.so = the previous build) on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "f32" -s -p no:cacheprovider > gpurun_out/s47_pytest_f32.log 2>&1
timeout 300 python scripts/f32_golden_report.py > gpurun_out/s47_f32_golden_digits.txt 2>&1
for f in standard-random,f32 f32; do
  n=$(echo $f | tr ',' '_')
  timeout 600 python bench.py --features $f --no-cpu-baseline > gpurun_out/s47_bench_${n}_direct.json 2> gpurun_out/s47_bench_${n}_direct.err
  TP3_LIB=$PWD/3photons-rust_b200/_build/libtp3_quadrant.so timeout 600 python bench.py --features $f --no-cpu-baseline > gpurun_out/s47_bench_${n}_quadrant.json 2> gpurun_out/s47_bench_${n}_quadrant.err
done
tail -15 gpurun_out/s47_pytest_f32.log | cut -c1-300
for f in gpurun_out/s47_bench_*.json; do python -c "
import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"; done
grep -n "selected" gpurun_out/s47_f32_golden_digits.txt
