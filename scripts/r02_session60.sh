#!/bin/bash
# session 60 (8 GPUs): bench lines of the final code at N = 8, default and f32 + xoshiro128+
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 10 --warmup 3 "${@:3}" > gpurun_out/s60_bench_$2.json 2> gpurun_out/s60_bench_$2.err; }
run 29511 n8
run 29512 f32xo_n8 --features standard-random,f32 --no-weak-subrecord
for f in n8 f32xo_n8; do python -c "
import json,sys; d=json.loads(open('gpurun_out/s60_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('weak'), d['clocks'])"; done
