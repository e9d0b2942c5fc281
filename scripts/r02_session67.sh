#!/bin/bash
# session 67 (2 GPUs): the single-process multi-device tests and the CLI with --gpus 2 on the final build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_device" -p no:cacheprovider -rA > gpurun_out/s67_pytest.log 2>&1
mkdir -p /tmp/run2 && cp tests/golden/valeurs /tmp/run2/ && (cd /tmp/run2 && timeout 120 /root/repo/3photons-rust_b200/_build/trois_photons_b200 --gpus 2 > stdout.log 2> stderr.log; echo "cli rc=$?"; python - <<'PY'
import sys
sys.path.insert(0, "/root/repo/tests")
from numdiff import compare
print("res.data vs golden at 1e-10:", compare(open("res.data").read(), open("/root/repo/tests/golden/res.data-features_").read(), rel=1e-10) or "identical within tolerance")
PY
) > gpurun_out/s67_cli.log 2>&1
grep -E "PASSED|FAILED|passed|failed" gpurun_out/s67_pytest.log | tail -6; cat gpurun_out/s67_cli.log
