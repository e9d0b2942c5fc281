#!/bin/bash
# session 43: the whole GPU suite and smoke() on the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/s43_pytest.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s43_smoke.log 2>&1
python scripts/ab_bits.py "" 20000 > gpurun_out/s43_bits.txt 2>&1
tail -3 gpurun_out/s43_pytest.log; tail -1 gpurun_out/s43_smoke.log; cat gpurun_out/s43_bits.txt
