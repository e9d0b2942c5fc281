#!/bin/bash
# session 70 (8 GPUs): second, independent N = 8 line of the final build (default config), and N = 1 on the same box for the ratio
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-weak-subrecord > gpurun_out/s70_bench_n8.json 2> gpurun_out/s70_bench_n8.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s70_bench_n1.json 2> gpurun_out/s70_bench_n1.err
for f in n8 n1; do python -c "
import json,sys; d=json.loads(open('gpurun_out/s70_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"; done
