#!/bin/bash
# session 46 (2 GPUs): multi-device tests incl. tp3_simulate_batches_merged on two devices; the CLI with --gpus 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_device" -p no:cacheprovider > gpurun_out/s46_pytest.log 2>&1
mkdir -p /tmp/run2 && cp tests/golden/valeurs /tmp/run2/ && (cd /tmp/run2 && timeout 120 /root/repo/3photons-rust_b200/_build/trois_photons_b200 --gpus 2 > stdout.log 2> stderr.log; echo "cli rc=$?"; python - <<'PY'
import sys
sys.path.insert(0, "/root/repo/tests")
from numdiff import compare
print("res.data vs golden at 1e-10:", compare(open("res.data").read(), open("/root/repo/tests/golden/res.data-features_").read(), rel=1e-10) or "identical within tolerance")
PY
) > gpurun_out/s46_cli.log 2>&1
tail -3 gpurun_out/s46_pytest.log; cat gpurun_out/s46_cli.log
