"""Print the SASS (with executed counts) attributed to given source lines.
usage: ncu_sass_for_lines.py report.ncu-rep lib.so kernel-substring file line_lo line_hi [max]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, lib, kern, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
mx = int(sys.argv[7]) if len(sys.argv) > 7 else 80
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
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; iA, iS, iI = h.index("Address"), h.index("Source"), h.index("Instructions Executed")
base = None; n = 0
for r in rows[2:]:
    if len(r) <= iI or not r[iI]: continue
    a = int(r[iA], 16) if r[iA].startswith("0x") else int(r[iA])
    if base is None: base = a
    l = loc.get(a - base)
    if l and l[0] == fname and lo <= l[1] <= hi and int(r[iI]) > 1000000:
        print(f"{a-base:06x} {l[1]:4d} {int(r[iI]):>10}  {r[iS].strip()}")
        n += 1
        if n >= mx: break
