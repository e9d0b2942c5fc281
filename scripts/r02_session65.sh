#!/bin/bash
# session 65: final build -- whole GPU suite, smoke(), default bench line + reference arm, launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/s65_pytest.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s65_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/s65_bench_n1.json 2> gpurun_out/s65_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s65_bench_ref.json 2> gpurun_out/s65_bench_ref.err
timeout 600 python bench.py --features standard-random,f32 --no-cpu-baseline > gpurun_out/s65_bench_f32xo.json 2> gpurun_out/s65_bench_f32xo.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s65_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s65_ncu_bench.log 2>&1
tail -3 gpurun_out/s65_pytest.log; tail -1 gpurun_out/s65_smoke.log
for f in n1 f32xo; do python -c "
import json,sys; d=json.loads(open('gpurun_out/s65_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"; done
tail -1 gpurun_out/s65_bench_ref.json | cut -c1-400
