#!/bin/bash
# session 50: log / sin cos tables from global memory with coalesced loads instead of divergent constant-bank loads: A/B against
# the previous build (libtp3_consttab.so) on the same box, bit comparison, then the tests that touch the tables
mkdir -p gpurun_out
B=3photons-rust_b200/_build
timeout 1200 python scripts/ab_libs_probe.py $B/libtp3_consttab.so $B/libtp3.so $B/libtp3_consttab.so $B/libtp3.so > gpurun_out/s50_ab.txt 2>&1
for f in "" "standard-random" "faster-evgen,no-photon-sorting" "f32"; do
  TP3_LIB=$PWD/$B/libtp3_consttab.so python scripts/ab_bits.py "$f" 4000 >> gpurun_out/s50_bits.txt 2>&1
  python scripts/ab_bits.py "$f" 4000 >> gpurun_out/s50_bits.txt 2>&1
done
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider -k "fastmath or events_match or golden or faster_evgen or histogram" > gpurun_out/s50_pytest.log 2>&1
cat gpurun_out/s50_ab.txt; cat gpurun_out/s50_bits.txt; tail -3 gpurun_out/s50_pytest.log
