#!/bin/bash
# session 24: batches cut into parts for small runs; context creation without cudaGetDeviceProperties
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "batch_parts or whole_program or cli_binary or default_run or stream_continues or histogram" -p no:cacheprovider > gpurun_out/s24_pytest.log 2>&1
timeout 300 python scripts/default_run_timing.py > gpurun_out/s24_default_run.txt 2>&1
timeout 300 python scripts/default_run_timing.py f32 > gpurun_out/s24_default_run_f32.txt 2>&1
tail -5 gpurun_out/s24_pytest.log; cat gpurun_out/s24_default_run.txt gpurun_out/s24_default_run_f32.txt
