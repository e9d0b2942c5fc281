#!/usr/bin/env python
"""faster-evgen stream pipeline: how should the walk of pass k + 1 and the physics of pass k share the SMs?
Sweeps the pass size (fe_pass_segments = 148 x 32 x k: the walk then holds k one-warp CTAs per SM and the physics kernel
the registers that are left) at 2e9 events.  usage: fe_corun_probe.py [features] [events]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.package()
features = sys.argv[1] if len(sys.argv) > 1 else "faster-evgen,no-photon-sorting"
n_events = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2 * 10**9
text = open(os.path.join(ROOT, "tests", "golden", "valeurs")).read()
cfg = pkg.Configuration.parse(text, features).with_num_events(n_events)
nb, last = pkg.batch_layout(n_events)
want = None
for k in (26, 20, 16, 13, 12, 10, 8, 6, 4):
    with pkg.Simulator(cfg) as sim:
        sim.set_option("fe_pass_segments", 148 * 32 * k)
        sim.simulate_merged(0, min(nb, 4000), 10000)  # warm-up: allocations, module load
        best = 1e30
        for _ in range(2):
            t0 = time.perf_counter()
            acc = sim.simulate_merged(0, nb, last)
            best = min(best, time.perf_counter() - t0)
        if want is None:
            want = bytes(acc)
        print(f"{features} {n_events:.0e} events, walk CTAs per SM {k:2d}: {best * 1e3:9.2f} ms  {n_events / best:.4g} events/s  "
              f"passes {sim.get_stat('fe_passes')}  same bits {bytes(acc) == want}", flush=True)
