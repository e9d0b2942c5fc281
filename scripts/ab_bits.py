"""Prints a digest of the merged accumulator and of a few per-batch accumulators for a feature set: run it with two builds
(TP3_LIB=...) and compare the lines.  usage: ab_bits.py features [n_batches]"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.package()
features = sys.argv[1] if len(sys.argv) > 1 else ""
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
text = open(os.path.join(g.ROOT, "tests", "golden", "valeurs")).read()
cfg = pkg.Configuration.parse(text, features)
with pkg.Simulator(cfg) as sim:
    accs = sim.simulate_batches(7, nb, 4321)
    merged = sim.simulate_merged(7, nb, 4321)
print(f"{features!r:45} per-batch sha1 {hashlib.sha1(bytes(accs)).hexdigest()[:16]}  merged sha1 {hashlib.sha1(bytes(merged)).hexdigest()[:16]}  selected {merged.selected_events} sigma {merged.sigma!r}")
