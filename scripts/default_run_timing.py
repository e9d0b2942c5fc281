"""Wall time of the whole default run (10^7 events) through tp3_run and its parts (development aid)."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
pkg = g.package()
text = open("tests/golden/valeurs").read()
d = tempfile.mkdtemp(); open(os.path.join(d, "valeurs"), "w").write(text)
for i in range(3):
    t0 = time.perf_counter(); out, secs = pkg.main_run(os.path.join(d, "valeurs"), d); t1 = time.perf_counter()
    print(f"tp3_run: wall {1e3*(t1-t0):.1f} ms, reference-style timed region {1e3*secs:.1f} ms")
cfg = pkg.Configuration.parse(text)
t0 = time.perf_counter(); sim = pkg.Simulator(cfg); t1 = time.perf_counter()
for i in range(3):
    t2 = time.perf_counter(); accs = sim.simulate_batches(0, 1000); t3 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms; simulate_batches(1000 batches) {1e3*(t3-t2):.2f} ms -> {1e7/(t3-t2):.3g} events/s")
sim.close(); shutil.rmtree(d)
