#!/bin/bash
# session 61 (4 GPUs): bench lines of the final code at N = 4 and N = 2
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $3 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $3 --steps 10 --warmup 3 "${@:4}" > gpurun_out/s61_bench_$2.json 2> gpurun_out/s61_bench_$2.err; }
run 29521 n4 4 --no-weak-subrecord
run 29522 n2 2 --no-weak-subrecord
for f in n4 n2; do python -c "
import json,sys; d=json.loads(open('gpurun_out/s61_bench_$f.json').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"; done
