#!/usr/bin/env python
"""When and where every unit of the fused kernel ran (diagnostic build `make -C 3photons-rust_b200 trace`, TP3_LIB is set here):
how far apart the warps are when the big units end, when the queue runs dry, and what the drain costs.
usage: unit_trace.py n_batches [option=value ...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["TP3_LIB"] = os.path.join(ROOT, "3photons-rust_b200", "_build", "libtp3_trace.so")
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

pkg = entry.package()
n = int(float(sys.argv[1]))
opts = [a.split("=") for a in sys.argv[2:]]
text = open(os.path.join(ROOT, "tests", "golden", "valeurs")).read()
lib = C.CDLL(os.environ["TP3_LIB"])
st = torch.cuda.current_stream()
out13 = torch.zeros(13, dtype=torch.float64, device="cuda")
trace = torch.zeros(3 * (n + 4096), dtype=torch.int64, device="cuda")
sim = pkg.Simulator(pkg.Configuration.parse(text, "").with_num_events(n * 10000))
sim.set_stream(st.cuda_stream)
for k, v in opts:
    sim.set_option(k, int(v))
for _ in range(2):
    sim.simulate_merged_device(0, n, 10000, out13.data_ptr())
torch.cuda.synchronize()
lib.tp3_debug_set_trace(C.c_void_p(trace.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
sim.simulate_merged_device(0, n, 10000, out13.data_ptr())
e1.record(st)
torch.cuda.synchronize()
lib.tp3_debug_set_trace(C.c_void_p(0))
t = trace.cpu().numpy().reshape(-1, 3)
used = t[:, 0] != 0
t = t[used]
start, end, meta = t[:, 0].astype(np.float64), t[:, 1].astype(np.float64), t[:, 2]
size = (meta >> 32).astype(np.int64)
smid, wslot = (meta & 0xFFFF).astype(np.int64), ((meta >> 16) & 0xFFFF).astype(np.int64)
t0 = start.min()
start, end = (start - t0) * 1e-6, (end - t0) * 1e-6  # ms
T = end.max()
W = 148 * 16
print(f"# {n} batches, options {opts}: {len(t)} units, kernel {e0.elapsed_time(e1):.3f} ms by events, {T:.3f} ms first start -> last end")
for sz in sorted(set(size.tolist()), reverse=True):
    m = size == sz
    d = end[m] - start[m]
    print(f"  units of {sz} batches: {m.sum():6d}   duration mean {d.mean():.3f} ms (per batch {d.mean() / sz:.4f}) sd {d.std():.3f}   first start {start[m].min():7.3f}  last start {start[m].max():7.3f}  last end {end[m].max():7.3f}")
big = size == size.max()
idx = np.nonzero(big)[0]
for w in range(0, int(np.ceil(big.sum() / W))):
    sel = idx[w * W:(w + 1) * W]
    dd = end[sel] - start[sel]
    print(f"  big units {w * W:6d}..{w * W + len(sel) - 1:6d}: duration mean {dd.mean():.3f} min {dd.min():.3f} max {dd.max():.3f}; start {start[sel].min():7.3f} .. {start[sel].max():7.3f} (spread {start[sel].max() - start[sel].min():.3f})   end {end[sel].min():7.3f} .. {end[sel].max():7.3f} (spread {end[sel].max() - end[sel].min():.3f})")
# batches completed per ms (device-wide), by the time their unit ends, in 2.5 ms windows
edges = np.arange(0.0, T + 2.5, 2.5)
hist, _ = np.histogram(end, bins=edges, weights=size.astype(np.float64))
print("  batches ended per ms, 2.5 ms windows: " + " ".join(f"{h / 2.5:.0f}" for h in hist))
for k in range(4):
    m = (wslot % 4) == k
    print(f"  warp slots = {k} mod 4: {m.sum()} units, mean duration per batch {((end[m] - start[m]) / size[m]).mean():.4f} ms")
# is the steady state a rigid pattern?  unit u and unit u + W on the same slot; sub-partitions of 592 consecutive units
nb = int(big.sum())
if nb > 5 * W:
    a, b = idx[3 * W:4 * W], idx[4 * W:5 * W]
    same = ((smid[a] == smid[b]) & (wslot[a] == wslot[b])).mean()
    print(f"  steady state: unit u and unit u + {W} run on the same (SM, warp slot) for {same:.3f} of the units of one wave")
    for g in range(4):
        sel = idx[3 * W + g * 592:3 * W + (g + 1) * 592]
        cnt = np.bincount(wslot[sel] % 4, minlength=4)
        per_sm = np.bincount(smid[sel], minlength=148)
        print(f"    units {3 * W + g * 592}..: warp slot mod 4 counts {cnt.tolist()}, units per SM min {per_sm.min()} max {per_sm.max()}, start {start[sel].min():.3f}..{start[sel].max():.3f}")
t_dry = start.max()
busy_after = np.clip(end - t_dry, 0, None).sum()  # slot-time still to run when the last unit starts
print(f"  last unit starts at {t_dry:.3f} ms; the drain takes {T - t_dry:.3f} ms with {busy_after / (T - t_dry) / W:.2f} of the slots busy on average")
print(f"  slot-time idle before the kernel ends: {(W * T - (end - start).sum()) / W:.3f} ms per slot")
# who finishes last: per sub-partition (warp slot % 4) and per SM
last = np.argsort(end)[-16:]
print("  the 16 units that end last: " + ", ".join(f"u{np.nonzero(used)[0][i]}(x{size[i]} sm{smid[i]} w{wslot[i]} {end[i]:.3f})" for i in last))
ends_by_sm = np.array([end[smid == s].max() for s in range(148)])
print(f"  last end per SM: min {ends_by_sm.min():.3f} median {np.median(ends_by_sm):.3f} max {ends_by_sm.max():.3f}")
for q in (0.5, 0.9, 0.99, 1.0):
    print(f"  {q:4.2f} of the slot-time is done by {np.interp(q, np.cumsum(np.sort(end) * 0 + (end - start)[np.argsort(end)]) / (end - start).sum(), np.sort(end)):.3f} ms", end=";")
print()
