#!/usr/bin/env python
"""faster-evgen stream pipeline: parts (2500 events) per CTA of the physics kernel.  `grid_warps` caps the number of CTAs of a pass
(about 77 600 parts per full pass), so 39 000 / 19 500 / 9 700 make a CTA take 2 / 4 / 8 consecutive parts.  Wall time of
tp3_simulate_merged through the C ABI at 2e9 events, best of 4.  usage: fe_units_probe.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.package()
text = open(os.path.join(ROOT, "tests", "golden", "valeurs")).read()
n_events = 2 * 10**9
cfg = pkg.Configuration.parse(text, "faster-evgen,no-photon-sorting").with_num_events(n_events)
nb, last = pkg.batch_layout(n_events)
for grid in (0, 39000, 19500, 9700, 0):
    with pkg.Simulator(cfg) as sim:
        sim.set_option("grid_warps", grid)
        sim.simulate_merged(0, 2000, 10000)
        best = 1e30
        for _ in range(4):
            t0 = time.perf_counter()
            acc = sim.simulate_merged(0, nb, last)
            best = min(best, time.perf_counter() - t0)
        print(f"grid_warps {grid:6d}: {best * 1e3:8.2f} ms  {n_events / best:.4g} events/s  selected {acc.selected_events} sigma {acc.sigma!r}", flush=True)
