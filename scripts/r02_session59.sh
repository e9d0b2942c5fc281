#!/bin/bash
# session 59: the whole GPU suite, smoke(), a unit trace and the bench lines of the final schedule (default, f32 + xoshiro, faster-evgen)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/s59_pytest.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s59_smoke.log 2>&1
timeout 200 python scripts/unit_trace.py 125000 > gpurun_out/s59_trace.txt 2>&1
timeout 200 python scripts/unit_trace.py 125000 taper_units=-1 align_units=0 >> gpurun_out/s59_trace.txt 2>&1
timeout 900 python bench.py > gpurun_out/s59_bench_n1.json 2> gpurun_out/s59_bench_n1.err
timeout 600 python bench.py --features standard-random,f32 --no-cpu-baseline > gpurun_out/s59_bench_f32xo.json 2> gpurun_out/s59_bench_f32xo.err
timeout 600 python bench.py --features f32 --no-cpu-baseline > gpurun_out/s59_bench_f32.json 2> gpurun_out/s59_bench_f32.err
timeout 600 python bench.py --features faster-evgen,no-photon-sorting --events 2e9 --no-cpu-baseline > gpurun_out/s59_bench_fe.json 2> gpurun_out/s59_bench_fe.err
tail -3 gpurun_out/s59_pytest.log; tail -1 gpurun_out/s59_smoke.log
for f in n1 f32xo f32 fe; do python -c "
import json,sys; d=json.loads(open('gpurun_out/s59_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"; done
