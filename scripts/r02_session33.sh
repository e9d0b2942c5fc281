#!/bin/bash
# session 33, 8 GPUs: final lines of the round (default strong scaling, faster-evgen tiles, f32 + xoshiro) and the multi-device tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_device" -p no:cacheprovider > gpurun_out/s33_pytest_multi.log 2>&1
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $R --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s33_bench_n8.json 2> gpurun_out/s33_bench_n8.err
timeout 600 $R --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 --features faster-evgen,no-photon-sorting > gpurun_out/s33_bench_fe_n8.json 2> gpurun_out/s33_bench_fe_n8.err
timeout 600 $R --master-port 29543 bench.py --gpus 8 --steps 10 --warmup 3 --features standard-random,f32 > gpurun_out/s33_bench_f32xo_n8.json 2> gpurun_out/s33_bench_f32xo_n8.err
R4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $R4 --master-port 29544 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/s33_bench_n4.json 2> gpurun_out/s33_bench_n4.err
tail -3 gpurun_out/s33_pytest_multi.log
python - <<'PY'
import json
for f in ['s33_bench_n8.json','s33_bench_fe_n8.json','s33_bench_f32xo_n8.json','s33_bench_n4.json']:
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, 'N=%d value %.4g e2e %.4g ms/step %.2f'%(d['n_gpus'],d['value'],d['e2e']['value'],d['ms_per_step']), d.get('weak'), d.get('check'), d.get('clocks'))
    except Exception as e: print(f, 'ERR', e)
PY
for f in gpurun_out/s33_*.err; do echo $f; tail -n 2 $f | cut -c1-200; done
