"""A/B timing of library variants: for each .so given, events/s of the default f64 RANF kernel (device time)."""
import os, subprocess, sys
code = r'''
import os, sys, time
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
pkg = g.package()
text = open("tests/golden/valeurs").read()
cfg = pkg.Configuration.parse(text, os.environ.get("TP3_FEATURES", ""))
nb = int(os.environ.get("TP3_NB", "200000"))
with pkg.Simulator(cfg) as sim:
    sim.simulate_batches_device(0, nb); sim.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); sim.simulate_batches_device(0, nb); sim.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"{os.environ.get('TP3_LIB','default'):60} {nb*1e4/best:.4g} events/s")
'''
for lib in sys.argv[1:]:
    env = dict(os.environ)
    if lib != "default": env["TP3_LIB"] = os.path.abspath(lib)
    subprocess.run([sys.executable, "-c", code], env=env)
