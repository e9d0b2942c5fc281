#!/bin/bash
# session 34 (2 GPUs): bench lines at N = 1 and N = 2 with the clock sampler started before the warm-up; multi-device tests; reference arm
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/s34_bench_n1.json 2> gpurun_out/s34_bench_n1.err
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $R --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s34_bench_n2.json 2> gpurun_out/s34_bench_n2.err
timeout 600 $R --master-port 29552 bench.py --gpus 2 --steps 5 --warmup 3 --features faster-evgen,no-photon-sorting --events 4e9 > gpurun_out/s34_bench_fe_n2.json 2> gpurun_out/s34_bench_fe_n2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s34_bench_ref.json 2> gpurun_out/s34_bench_ref.err
python - <<'PY'
import json
for f in ['s34_bench_n1.json','s34_bench_n2.json','s34_bench_fe_n2.json','s34_bench_ref.json']:
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, 'N=%d value %.4g e2e %.4g ms/step %.2f'%(d['n_gpus'],d['value'],d['e2e']['value'],d['ms_per_step']), d.get('clocks'), d.get('e2e_per_batch'))
    except Exception as e: print(f, 'ERR', e)
PY
