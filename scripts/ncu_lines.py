"""Attribute the per-instruction counts of an ncu report to CUDA source lines.
usage: ncu_lines.py report.ncu-rep lib.so kernel-substring n_events [top]
Needs the SAME lib.so that was profiled (addresses are matched against nvdisasm -g output)."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, lib, kern, n_events = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "host" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# address -> (file, line) for the target function
loc = {}; cur = None; infn = False
for line in dis.splitlines():
    if line.startswith(".text."):
        infn = kern in line
        continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m: loc[int(m.group(1), 16)] = (cur, m.group(2).strip())
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; iA, iS, iI = h.index("Address"), h.index("Source"), h.index("Instructions Executed")
base = None
byline = collections.Counter(); byfile = collections.Counter(); fp64 = collections.Counter(); tot = 0; miss = 0
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
for r in rows[2:]:
    if len(r) <= iI or not r[iI]: continue
    a = int(r[iA], 16) if r[iA].startswith("0x") else int(r[iA])
    if base is None: base = a
    n = int(r[iI]); tot += n
    l = loc.get(a - base)
    if l is None: miss += n; continue
    key = l[0]
    byline[key] += n; byfile[key[0] if key else "?"] += n
    op = re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip()).split(".")[0].split()[0]
    if op in FP64: fp64[key] += n
print(f"total {tot*32/n_events:.1f} slots/event (unmatched {miss*32/n_events:.1f})")
for f, n in byfile.most_common(): print(f"  {f:22} {n*32/n_events:8.1f}")
srcs = {}
for key, n in byline.most_common(top):
    f, ln = key
    if f not in srcs:
        path = os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc", f)
        srcs[f] = open(path).read().splitlines() if os.path.exists(path) else []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    print(f"{n*32/n_events:7.1f} (fp64 {fp64[key]*32/n_events:6.1f})  {f}:{ln}  {text}")
