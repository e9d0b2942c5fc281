"""Stall-sample attribution by source line. usage: ncu_stalls.py report lib.so kernel-substring [top]"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "host" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
loc = {}; cur = None; infn = False
for line in dis.splitlines():
    if line.startswith(".text."): infn = kern in line; continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m: loc[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); h = rows[1]
iA, iI, iN = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
base = None; samp = collections.Counter(); inst = collections.Counter(); stl = collections.defaultdict(collections.Counter)
byfile = collections.Counter()
for r in rows[2:]:
    if len(r) <= iN or not r[iI]: continue
    a = int(r[iA], 16) if r[iA].startswith("0x") else int(r[iA])
    if base is None: base = a
    l = loc.get(a - base)
    samp[l] += int(r[iN] or 0); inst[l] += int(r[iI]); byfile[l[0] if l else "?"] += int(r[iN] or 0)
    for i in stall_cols:
        if r[i]: stl[l][h[i]] += int(r[i])
tot = sum(samp.values()); ti = sum(inst.values())
agg = collections.Counter()
for l in stl:
    for k, v in stl[l].items(): agg[k] += v
print("stall mix:", {k[6:]: round(100 * v / tot, 1) for k, v in agg.most_common(9)})
print("samples by file:", {k: round(100 * v / tot, 1) for k, v in byfile.most_common()})
for l, v in samp.most_common(top):
    t3 = ", ".join(f"{k[6:]}:{100*x/v:.0f}%" for k, x in stl[l].most_common(3))
    print(f"{100*v/tot:5.1f}% samples {100*inst[l]/ti:5.1f}% inst  {l}  [{t3}]")
