#!/bin/bash
# session 48: the whole GPU suite on the build with the single-call f32 sin / cos, smoke(), ncu --set full of the packed f32 kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/s48_pytest.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s48_smoke.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simulate_kernel_x2 -c 1 -o gpurun_out/s48_f32x2_kernel python scripts/ncu_target.py 200000 standard-random,f32 0 1 > gpurun_out/s48_ncu.log 2>&1
tail -3 gpurun_out/s48_pytest.log; tail -1 gpurun_out/s48_smoke.log; tail -2 gpurun_out/s48_ncu.log
