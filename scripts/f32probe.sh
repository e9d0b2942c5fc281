TP3_FEATURES=f32 python scripts/ab_probe.py default 2>&1 | grep events
TP3_FEATURES=standard-random,f32 python scripts/ab_probe.py default 2>&1 | grep events
TP3_FEATURES=standard-random python scripts/ab_probe.py default 2>&1 | grep events
python -m pytest tests -q -m gpu -x -k "f32" 2>&1 | tail -3
