#!/bin/bash
# session 35: xoshiro256+ f64 generation straight from the stream integers (exact rescalings): same bits as before, fewer FP64 instructions
mkdir -p gpurun_out
for f in "standard-random" "standard-random,multi-threading,faster-threading" "standard-random,no-photon-sorting" "" "standard-random,f32"; do
  TP3_LIB=$PWD/3photons-rust_b200/_build/libtp3_prev.so python scripts/ab_bits.py "$f" >> gpurun_out/s35_bits_prev.txt 2>&1
  python scripts/ab_bits.py "$f" >> gpurun_out/s35_bits_new.txt 2>&1
done
timeout 1500 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/s35_pytest.log 2>&1
timeout 600 python bench.py --features standard-random --no-cpu-baseline > gpurun_out/s35_bench_xo.json 2> gpurun_out/s35_bench_xo.err
timeout 600 python bench.py --features standard-random,multi-threading,faster-threading --no-cpu-baseline > gpurun_out/s35_bench_xo_ft.json 2> gpurun_out/s35_bench_xo_ft.err
diff gpurun_out/s35_bits_prev.txt gpurun_out/s35_bits_new.txt && echo "SAME BITS"; cat gpurun_out/s35_bits_new.txt; tail -3 gpurun_out/s35_pytest.log
for f in xo xo_ft; do python -c "
import json,sys; d=json.loads(open('gpurun_out/s35_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"; done
