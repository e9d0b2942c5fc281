"""Quick GPU-side timing probe (development aid, not the bench): events/s of each kernel variant."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.package()
text = open(os.path.join(g.ROOT, "tests", "golden", "valeurs")).read()
n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for features, kernel in [("", 0), ("", 1), ("no-photon-sorting", 0), ("standard-random", 0), ("f32", 0), ("standard-random,f32", 0)]:
    cfg = pkg.Configuration.parse(text, features)
    with pkg.Simulator(cfg, kernel) as sim:
        if features == "" and kernel == 0:
            print("peak fp64 TFLOP/s", sim.peak_probe(0), "fp32", sim.peak_probe(1), flush=True)
        sim.simulate_merged(0, 2000)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            acc = sim.simulate_merged(0, n_batches)
            best = min(best, time.perf_counter() - t0)
        ev = n_batches * 10000
        print(f"features={features!r:28} kernel={kernel} {ev/best:.4g} events/s ({best*1e3:.1f} ms) selected_frac={acc.selected_events/ev:.6f}", flush=True)
# per-event observable epilogue (tp3_histograms_enable): cost relative to the plain kernel
cfg = pkg.Configuration.parse(text, "")
with pkg.Simulator(cfg, 0) as sim:
    for bins in (0, 200, 1024):
        sim.histograms_enable(bins)
        sim.simulate_merged(0, 2000)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            sim.simulate_merged(0, n_batches)
            best = min(best, time.perf_counter() - t0)
        print(f"histograms bins={bins:5d} {n_batches * 10000 / best:.4g} events/s", flush=True)
