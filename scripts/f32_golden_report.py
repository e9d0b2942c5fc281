#!/usr/bin/env python
"""f32 runs of the default `valeurs` on the GPU against the reference's f32 goldens, token by token.

For each f32 feature set x kernel (fast packed, fast one-event-per-lane, literal) prints, per line of res.data, the
largest absolute and relative difference of its numeric tokens (the 10 spin rows included) and whether the reference
CI's own bars hold (ci.yml:179-203: res.data abs 1.1e-8, stdout rel 1.9e-5).  Output kept as profiles/r02_f32_golden_digits.txt.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
from numdiff import _NUM  # noqa: E402


def token_diffs(actual, expected):
    rows = []
    for ln, (la, le) in enumerate(zip(actual.strip().splitlines(), expected.strip().splitlines()), 1):
        ta, te = la.split(), le.split()
        worst_abs = worst_rel = 0.0
        n = 0
        for a, e in zip(ta, te):
            if _NUM.match(a) and _NUM.match(e):
                fa, fe = float(a), float(e)
                if fa != fa and fe != fe:
                    continue
                n += 1
                d = abs(fa - fe)
                worst_abs = max(worst_abs, d)
                if max(abs(fa), abs(fe)) > 0:
                    worst_rel = max(worst_rel, d / max(abs(fa), abs(fe)))
        if n:
            rows.append((ln, le.strip()[:34], worst_abs, worst_rel, la.strip() == le.strip()))
    return rows


def main():
    pkg = entry.package()
    valeurs = open(os.path.join(ROOT, "tests", "golden", "valeurs")).read()
    for features in ("f32", "standard-random,f32"):
        want_rd = open(os.path.join(ROOT, "tests", "golden", "res.data-features_" + features)).read()
        want_so = open(os.path.join(ROOT, "tests", "golden", "stdout.log-features_" + features)).read()
        for name, kernel, scalar in (("fast (packed x2)", pkg.KERNEL_FAST, 0), ("fast (one event per lane)", pkg.KERNEL_FAST, 1),
                                     ("literal", pkg.KERNEL_LITERAL, 0)):
            cfg = pkg.Configuration.parse(valeurs, features)
            nb, last = pkg.batch_layout(cfg.num_events)
            with pkg.Simulator(cfg, kernel) as sim:
                sim.set_option("f32_scalar", scalar)
                fin = pkg.finalize(cfg, sim.simulate_merged(0, nb, last))
            print(f"=== features {features!r}, kernel {name}: selected {fin.selected_events}")
            rows = token_diffs(fin.res_data(), want_rd)
            ci_ok = all(r[2] <= 1.1e-8 for r in rows)
            for ln, text, wa, wr, same in rows:
                print(f"  res.data line {ln:2d} {text:<34s} max|d|={wa:9.3g} max rel={wr:9.3g} {'identical' if same else ''}")
            print(f"  res.data within CI's abs 1.1e-8: {ci_ok}; lines printed identically: {sum(r[4] for r in rows)}/{len(rows)}")
            srows = token_diffs(fin.stdout(), want_so)
            print(f"  stdout: max rel {max(r[3] for r in srows):.3g} (CI: 1.9e-5): {all(r[3] <= 1.9e-5 for r in srows)}")


if __name__ == "__main__":
    main()
