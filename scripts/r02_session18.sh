#!/bin/bash
# 8 GPUs: the strong-scaling line (BASELINE configs[4] as written), faster-evgen stream tiles, xoshiro faster-threading, the CPU arm
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $R --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s18_bench_n8.json 2> gpurun_out/s18_bench_n8.err
timeout 600 $R --master-port 29532 bench.py --gpus 8 --steps 5 --warmup 3 --features faster-evgen,no-photon-sorting > gpurun_out/s18_bench_fe_n8.json 2> gpurun_out/s18_bench_fe_n8.err
timeout 600 $R --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --features standard-random,multi-threading,faster-threading --no-weak-subrecord > gpurun_out/s18_bench_xo_ft_n8.json 2> gpurun_out/s18_bench_xo_ft_n8.err
timeout 600 python bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/s18_bench_ref_n8.json 2> gpurun_out/s18_bench_ref_n8.err
python - <<'PY'
import json
for f in ['s18_bench_n8.json','s18_bench_fe_n8.json','s18_bench_xo_ft_n8.json','s18_bench_ref_n8.json']:
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, 'N=%d value %.4g e2e %.4g ms/step %.2f'%(d['n_gpus'],d['value'],d['e2e']['value'],d['ms_per_step']), d.get('weak'), d.get('check'), d.get('clocks'))
    except Exception as e: print(f, 'ERR', e)
PY
for f in gpurun_out/s18_*.err; do echo $f; tail -n 2 $f | cut -c1-200; done
