"""Small launches of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.package()
text = open(os.path.join(g.ROOT, "tests", "golden", "valeurs")).read()
for features, kernel in [("", 0), ("", 1), ("standard-random", 0), ("f32", 0), ("multi-threading,faster-threading", 0), ("faster-evgen", 0),
                         ("faster-evgen,standard-random", 0)]:
    cfg = pkg.Configuration.parse(text, features).with_num_events(25000)
    with pkg.Simulator(cfg, kernel) as sim:
        accs = sim.simulate_batches(0, 3, 5000)
        m = sim.simulate_merged(0, 3, 5000)
        if "faster-evgen" not in features:
            sim.rng_dump(1, 2400)
        print(features or "default", kernel, [a.selected_events for a in accs], m.selected_events, flush=True)
# faster-evgen with one thread per batch (the scan supplies batch starts only) and the per-event observable epilogue
os.environ["TP3_FE_SPLIT"] = "1"
cfg = pkg.Configuration.parse(text, "faster-evgen,f32").with_num_events(25000)
with pkg.Simulator(cfg, 0) as sim:
    print("faster-evgen,f32 split 1", [a.selected_events for a in sim.simulate_batches(0, 3, 5000)], flush=True)
cfg = pkg.Configuration.parse(text, "faster-evgen,standard-random,f32").with_num_events(25000)
with pkg.Simulator(cfg, 0) as sim:  # xoshiro scan (fe_scan_xo.cuh), batch starts only
    print("faster-evgen,standard-random,f32 split 1", [a.selected_events for a in sim.simulate_batches(0, 3, 5000)], flush=True)
del os.environ["TP3_FE_SPLIT"]
cfg = pkg.Configuration.parse(text, "").with_num_events(25000)
with pkg.Simulator(cfg, 0) as sim:
    sim.histograms_enable(200)
    sim.simulate_batches(0, 3, 5000)
    print("histograms", sum(sim.histograms_fetch().counts[0]), flush=True)
