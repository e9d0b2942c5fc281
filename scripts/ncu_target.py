"""Minimal launch sequence for ncu: the fused kernel on `n_batches` batches of the default configuration.
usage: ncu_target.py [n_batches] [features] [kernel] [merged]   (merged = 1: with the in-kernel ordered fold, as bench.py launches it)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.package()
text = open(os.path.join(g.ROOT, "tests", "golden", "valeurs")).read()
n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 2960
features = sys.argv[2] if len(sys.argv) > 2 else ""
kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
merged = len(sys.argv) > 4 and sys.argv[4] == "1"
cfg = pkg.Configuration.parse(text, features)
with pkg.Simulator(cfg, kernel) as sim:
    for _ in range(3):
        if merged:
            sim.simulate_merged(0, n_batches)
        else:
            sim.simulate_batches_device(0, n_batches)
            sim.synchronize()
    print("done", sim.launch_count)
