#!/usr/bin/env python
"""faster-evgen stream pipeline: segment length (RANF rounds per lane and pass) against the rate.  usage: fe_seg_probe.py [events]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
pkg = entry.package()
features = "faster-evgen,no-photon-sorting"
n_events = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4 * 10**9
text = open(os.path.join(ROOT, "tests", "golden", "valeurs")).read()
cfg = pkg.Configuration.parse(text, features).with_num_events(n_events)
nb, last = pkg.batch_layout(n_events)
want = None
for seg in (0, 256, 512, 1024):
    with pkg.Simulator(cfg) as sim:
        sim.set_option("fe_seg_rounds", seg)
        sim.simulate_merged(0, min(nb, 4000), 10000)
        best = 1e30
        for _ in range(2):
            t0 = time.perf_counter()
            acc = sim.simulate_merged(0, nb, last)
            best = min(best, time.perf_counter() - t0)
        want = want or bytes(acc)
        print(f"{n_events:.0e} events, fe_seg_rounds {seg:5d}: {best * 1e3:9.2f} ms  {n_events / best:.4g} events/s  passes {sim.get_stat('fe_passes')}  same bits {bytes(acc) == want}", flush=True)
