#!/bin/bash
# session 37: RANF r r' as an exact 64-bit integer product converted once: same bits?  faster?
mkdir -p gpurun_out
rm -f gpurun_out/s37_bits_*.txt
for f in "" "no-photon-sorting" "standard-random" "faster-evgen" "faster-evgen,standard-random"; do
  TP3_LIB=$PWD/3photons-rust_b200/_build/libtp3_prev.so python scripts/ab_bits.py "$f" 20000 >> gpurun_out/s37_bits_prev.txt 2>&1
  python scripts/ab_bits.py "$f" 20000 >> gpurun_out/s37_bits_new.txt 2>&1
done
diff gpurun_out/s37_bits_prev.txt gpurun_out/s37_bits_new.txt && echo "SAME BITS"; cat gpurun_out/s37_bits_new.txt
TP3_LIB=$PWD/3photons-rust_b200/_build/libtp3_prev.so timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s37_bench_prev.json 2> gpurun_out/s37_bench_prev.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s37_bench_new.json 2> gpurun_out/s37_bench_new.err
TP3_LIB=$PWD/3photons-rust_b200/_build/libtp3_prev.so timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s37_bench_prev2.json 2> gpurun_out/s37_bench_prev2.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s37_bench_new2.json 2> gpurun_out/s37_bench_new2.err
timeout 600 python bench.py --no-cpu-baseline --features faster-evgen,no-photon-sorting --events 2e9 > gpurun_out/s37_bench_fe_new.json 2> gpurun_out/s37_bench_fe_new.err
for f in prev new prev2 new2 fe_new; do python -c "
import json,sys; d=json.loads(open('gpurun_out/s37_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['check']['selected_events'])"; done
