#!/usr/bin/env python
"""The 'no-cuts' configuration (valeurs:26-29, every cut off): how far are the GPU kernels from the oracle per batch, and why?
Prints, for the fast and the literal kernel, the largest relative deviation of each per-batch sum over 6 batches, and for batch 0
the share of the sums carried by the single largest event (collinear photons make the matrix elements singular).
Output kept as profiles/r02_nocuts.txt."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import oracle_lib  # noqa: E402
from test_gpu_parity import _edit_valeurs, acc_fields  # noqa: E402

pkg = entry.package()
text = open(os.path.join(ROOT, "tests", "golden", "valeurs")).read()
nb = 6
text = _edit_valeurs(text, i0=str(nb * 10000), i2="1.e0", i3="1.e0", i4="0.e0", i5="0.e0")
run = oracle_lib.run(text, "", want_text=False)
cfg = pkg.Configuration.parse(text)
names = ["spm2[A]", "spm2[B+]", "spm2[B-]", "spm2[R_MX]", "spm2[I_MX]", "vars[A]", "vars[B+]", "vars[B-]", "vars[R_MX]", "vars[I_MX]", "sigma", "variance"]
for kernel, kname in ((0, "fast"), (1, "literal")):
    with pkg.Simulator(cfg, kernel) as sim:
        accs = sim.simulate_batches(0, nb)
        mom, kept, m2 = sim.events_dump(0, 10000)
    worst = [0.0] * 12
    for b in range(nb):
        g, w = acc_fields(accs[b]), acc_fields(run.per_batch[b])
        for k in range(12):
            scale = abs(w[k]) if k >= 5 else max(abs(w[k]), math.sqrt(max(w[5 + k], 0.0)))
            worst[k] = max(worst[k], abs(g[k] - w[k]) / scale)
    print(f"kernel {kname}: largest relative deviation from the oracle over {nb} batches of 10000 events (all selected)")
    print("   " + "  ".join(f"{n} {v:.1e}" for n, v in zip(names, worst)))
    omom, okept, om2 = oracle_lib.events(text, "", 10000)
    # batch 0: the events that carry the sums, and how well they agree one by one
    tot = [sum(om2[e * 5 + k] for e in range(10000)) for k in range(3)]
    big = sorted(range(10000), key=lambda e: -om2[e * 5])[:3]
    for e in big:
        rel = [abs(m2[e * 5 + k] - om2[e * 5 + k]) / abs(om2[e * 5 + k]) for k in range(3)]
        cosb = max(abs(omom[(e * 3 + q) * 4 + 0]) / omom[(e * 3 + q) * 4 + 3] for q in range(3))
        print(f"   event {e}: {om2[e * 5] / tot[0]:.1%} of the batch's A sum, max |cos(photon, beam)| = 1 - {1 - cosb:.2e}, "
              f"GPU vs oracle on this event: A {rel[0]:.1e}  B+ {rel[1]:.1e}  B- {rel[2]:.1e}")
