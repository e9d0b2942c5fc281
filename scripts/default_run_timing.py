"""Wall time of the whole default run (10^7 events, BASELINE configs[1]) through tp3_run, stage by stage
(the reference's own timed region is main.rs:83-85,138 = context creation + simulation + finalize).
usage: default_run_timing.py [features]"""
import os, sys, time, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.package()
features = sys.argv[1] if len(sys.argv) > 1 else ""
text = open(os.path.join(ROOT, "tests", "golden", "valeurs")).read()
d = tempfile.mkdtemp()
open(os.path.join(d, "valeurs"), "w").write(text)
print(f"# tp3_run_stages, default valeurs (1e7 events), features {features!r}; run 0 includes CUDA initialisation and module load")
for i in range(4):
    t0 = time.perf_counter()
    out, secs, stages = pkg.main_run_stages(os.path.join(d, "valeurs"), d, features)
    wall = time.perf_counter() - t0
    print(f"run {i}: wall {1e3 * wall:8.2f} ms, reference-style timed region {1e3 * secs:8.2f} ms  |  " +
          "  ".join(f"{k} {1e3 * v:.2f} ms" for k, v in stages.items()))
cfg = pkg.Configuration.parse(text, features)
for parts in (1, 2, 5, 10, 0):
    with pkg.Simulator(cfg) as sim:
        sim.set_option("batch_parts", parts)
        best = 1e30
        for i in range(6):
            t0 = time.perf_counter()
            acc = sim.simulate_merged(0, 1000)
            best = min(best, time.perf_counter() - t0)
        print(f"tp3_simulate_merged(1000 batches) on a warm context, batch_parts = {parts}: {1e3 * best:.3f} ms -> {1e7 / best:.3g} events/s  (selected {acc.selected_events}, sigma sum {acc.sigma!r})")
shutil.rmtree(d)
