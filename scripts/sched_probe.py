#!/usr/bin/env python
"""A/B of launch schedules on one B200 (device time, CUDA events): the round-1 library (one CTA per 1-8 batches, many
waves) against the static balanced schedule of round 2, with and without the in-kernel ordered fold, for several
unit sizes.  usage: sched_probe.py [n_batches] [features]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
n_batches = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
features = sys.argv[2] if len(sys.argv) > 2 else ""

CHILD = r'''
import ctypes as C, os, sys, json
sys.path.insert(0, %(root)r)
import torch
import __graft_entry__ as entry
pkg = entry.package()
n, features, mode, unit, grid = %(n)d, %(features)r, %(mode)r, %(unit)d, %(grid)d
text = open(os.path.join(%(root)r, "tests", "golden", "valeurs")).read()
cfg = pkg.Configuration.parse(text, features).with_num_events(n * 10000)
sim = pkg.Simulator(cfg)
st = torch.cuda.current_stream()
sim.set_stream(st.cuda_stream)
if unit >= 0 and hasattr(sim, "set_option"):
    sim.set_option("unit_batches", unit)
if grid > 0:
    sim.set_option("grid_warps", grid)
out13 = torch.zeros(13, dtype=torch.float64, device="cuda")
def step():
    if mode == "nofold":
        sim.simulate_batches_device(0, n)
    else:
        sim.simulate_merged_device(0, n, 10000, out13.data_ptr())
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
K = 5
for _ in range(K): step()
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"ms": round(ms, 3), "events_per_s": float("%%.4g" %% (n * 10000 / (ms * 1e-3)))}))
'''


def run(lib, mode, unit, grid=0):
    env = dict(os.environ)
    if lib:
        env["TP3_LIB"] = lib
    code = CHILD % {"root": ROOT, "n": n_batches, "features": features, "mode": mode, "unit": unit, "grid": grid}
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    return r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else "FAILED " + r.stderr[-300:]


print(f"# {n_batches} batches, features {features!r}", flush=True)
for unit, grid in ((-1, 0), (8, n_batches // 8), (4, n_batches // 4), (1, n_batches), (8, 2368 * 4), (8, 2368 * 16), (1, 0)):
    print(f"unit_batches={unit:3d} grid_warps={grid:8d} (0 = resident warps, static), no fold        :", run(None, "nofold", unit, grid), flush=True)
    print(f"unit_batches={unit:3d} grid_warps={grid:8d} (0 = resident warps, static), in-kernel fold :", run(None, "fold", unit, grid), flush=True)
