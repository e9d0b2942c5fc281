#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/fe_probe2.py > gpurun_out/s15_fe_probe.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "faster_evgen or full_size or device_merge or ragged" -p no:cacheprovider > gpurun_out/s15_pytest.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s15_bench_n1.json 2> gpurun_out/s15_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s15_fe_launches.csv python scripts/ncu_target.py 20000 faster-evgen,no-photon-sorting > gpurun_out/s15_ncu1.log 2>&1
cat gpurun_out/s15_fe_probe.txt; tail -5 gpurun_out/s15_pytest.log; tail -2 gpurun_out/s15_bench.err; grep fe_ gpurun_out/s15_fe_launches.csv | tail -3 | cut -c60-250
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s15_bench_n1.json').read())
print('value %.4g e2e %.4g per_batch %.4g'%(d['value'],d['e2e']['value'],d['e2e_per_batch']['value']), d['check'])
PY
