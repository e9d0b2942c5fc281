"""Aggregate an `ncu --page source --csv` dump: warp-level instructions executed per opcode, per event.
usage: ncu_mix.py src.csv n_events [top]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
n_events = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
hdr = rows[1]; iS = hdr.index("Source"); iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed")
iSamp = hdr.index("# Samples")
by = collections.Counter(); byt = collections.Counter(); samp = collections.Counter(); tot = 0; tott = 0
for r in rows[2:]:
    if len(r) <= iT or not r[iI]: continue
    s = r[iS].strip(); s = re.sub(r"^@!?U?P\d+\s+", "", s)
    op = s.split()[0] if s else "?"
    fam = op.split(".")[0]
    key = op if fam in ("MUFU", "I2F", "F2I", "F2F", "DSETP", "DMNMX") else fam
    n = int(r[iI]); t = int(r[iT])
    by[key] += n; byt[key] += t; tot += n; tott += t; samp[key] += int(r[iSamp] or 0)
print(f"total warp-inst {tot:.4g}  = {tot*32/n_events:.1f} lane-slots/event ; thread-inst/event {tott/n_events:.1f}")
ts = sum(samp.values())
for k, n in by.most_common(top):
    print(f"{k:22} {n*32/n_events:8.1f} slots/ev {100*n/tot:5.1f}%  thr/ev {byt[k]/n_events:8.1f}  stall-samples {100*samp[k]/ts:5.1f}%")
