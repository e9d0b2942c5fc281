#!/usr/bin/env python
"""faster-evgen (sequential RANF stream) through the C ABI: events/s (wall clock around tp3_simulate_merged) of the
stream pipeline (fe_stream.cuh) and of the round-1 pipeline (`fe_legacy`), at a few run sizes.
usage: fe_probe2.py [features]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.package()
features = sys.argv[1] if len(sys.argv) > 1 else "faster-evgen,no-photon-sorting"
text = open(os.path.join(ROOT, "tests", "golden", "valeurs")).read()
for n_events in (10**7, 2 * 10**8, 2 * 10**9):
    cfg = pkg.Configuration.parse(text, features).with_num_events(n_events)
    nb, last = pkg.batch_layout(n_events)
    for legacy, serial in ((0, 0), (0, 1), (1, 0)):
        with pkg.Simulator(cfg) as sim:
            sim.set_option("fe_legacy", legacy).set_option("fe_serial", serial)
            sim.simulate_merged(0, min(nb, 2000), 10000)  # warm-up: allocations, module load
            best = 1e30
            for _ in range(3):
                t0 = time.perf_counter()
                acc = sim.simulate_merged(0, nb, last)
                best = min(best, time.perf_counter() - t0)
            extra = "" if legacy else f" passes {sim.get_stat('fe_passes')} redone segments {sim.get_stat('fe_redone')}"
            print(f"{features} {n_events:.0e} events, {'round-1 pipeline' if legacy else 'stream pipeline, passes one after the other' if serial else 'stream pipeline, walk next to physics    '}: {best * 1e3:9.2f} ms  "
                  f"{n_events / best:.4g} events/s  selected {acc.selected_events}{extra}", flush=True)
