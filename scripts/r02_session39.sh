#!/bin/bash
# session 39: final evidence of the round: whole GPU suite, ncu of the bench's own launch (default f64), launch list of bench.py, bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/s39_pytest.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -c 1 -o gpurun_out/s39_bench_kernel python scripts/ncu_target.py 1000000 "" 0 1 > gpurun_out/s39_ncu_kernel.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s39_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s39_ncu_bench.log 2>&1
timeout 900 python bench.py > gpurun_out/s39_bench_n1.json 2> gpurun_out/s39_bench_n1.err
timeout 600 python bench.py --features faster-evgen,no-photon-sorting --events 2e9 --no-cpu-baseline > gpurun_out/s39_bench_fe_n1.json 2> gpurun_out/s39_bench_fe_n1.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s39_smoke.log 2>&1
tail -3 gpurun_out/s39_pytest.log; tail -2 gpurun_out/s39_smoke.log
for f in n1 fe_n1; do python -c "
import json,sys; d=json.loads(open('gpurun_out/s39_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('frac_executed'), d['clocks'])"; done
