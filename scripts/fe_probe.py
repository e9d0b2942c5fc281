"""faster-evgen timing probe (development aid): events/s of the exact sequential-stream path at several sizes,
the ROUND-1 pipeline (`fe_legacy`): one thread per batch (fe_split 1) vs one lane per 313 events (fe_split 32); the option
`fe_timing` prints the scan phases.  The shipped stream pipeline is timed by scripts/fe_probe2.py."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.package()
text = open(os.path.join(g.ROOT, "tests", "golden", "valeurs")).read()
sizes = [int(x) for x in sys.argv[1:]] or [1000, 20000, 200000]
for features in ["faster-evgen,no-photon-sorting", "faster-evgen,f32"]:
    cfg = pkg.Configuration.parse(text, features)
    with pkg.Simulator(cfg, 0) as sim:
        sim.set_option("fe_legacy", 1)
        sim.simulate_merged(0, 100)
        for n_batches in sizes:
            for split in ("1", "32"):
                sim.set_option("fe_split", int(split))
                best = 1e9
                for _ in range(2):
                    t0 = time.perf_counter()
                    acc = sim.simulate_merged(0, n_batches)
                    best = min(best, time.perf_counter() - t0)
                ev = n_batches * 10000
                print(f"features={features!r:34} batches={n_batches:8d} split={split:>2} {ev/best:.4g} events/s ({best*1e3:.1f} ms) selected={acc.selected_events}", flush=True)
