import os, sys
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
pkg = g.package()
text = open("tests/golden/valeurs").read()
cfg = pkg.Configuration.parse(text, "faster-evgen")
with pkg.Simulator(cfg) as sim:
    sim.simulate_batches_device(0, 100000); sim.synchronize()
