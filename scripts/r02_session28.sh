#!/bin/bash
# session 28: tp3_simulate_batches_merged (tests + the bench's per-batch figures)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "streamed or one_call or batch_range or bad_arguments" -p no:cacheprovider > gpurun_out/s28_pytest.log 2>&1
timeout 900 python bench.py > gpurun_out/s28_bench_n1.json 2> gpurun_out/s28_bench_n1.err
tail -3 gpurun_out/s28_pytest.log; python -c "
import json; d=json.loads(open('gpurun_out/s28_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['e2e_per_batch']['value'], d['e2e_per_batch']['host_fold_value'], d['clocks'])"
