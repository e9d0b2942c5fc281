#!/bin/bash
# 2 GPUs: the single-process multi-device tests, and the N = 2 bench lines (default strong, faster-evgen stream tiles)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/s17_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_device or streamed" -p no:cacheprovider -rs > gpurun_out/s17_pytest_multidev.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s17_bench_n2.json 2> gpurun_out/s17_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 5 --warmup 3 --features faster-evgen,no-photon-sorting --events 4e9 > gpurun_out/s17_bench_fe_n2.json 2> gpurun_out/s17_bench_fe_n2.err
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --features faster-evgen,no-photon-sorting --events 4e9 --no-cpu-baseline > gpurun_out/s17_bench_fe_n1.json 2> gpurun_out/s17_bench_fe_n1.err
tail -5 gpurun_out/s17_pytest_multidev.log; tail -3 gpurun_out/s17_bench_n2.err gpurun_out/s17_bench_fe_n2.err
python - <<'PY'
import json
for f in ['s17_bench_n2.json','s17_bench_fe_n2.json','s17_bench_fe_n1.json']:
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, 'N=%d value %.4g e2e %.4g ms/step %.2f'%(d['n_gpus'],d['value'],d['e2e']['value'],d['ms_per_step']), d.get('weak'), d['check'])
    except Exception as e: print(f, 'ERR', e)
PY
