#!/usr/bin/env python
"""A/B of two builds of libtp3.so on one box (child processes, TP3_LIB): device time of the default fused kernel (in-kernel fold) at
the launch sizes of the 1- and 8-GPU shares of the 1e10-event run, and wall time of faster-evgen through the C ABI at 2e9 events.
usage: ab_libs_probe.py libA.so libB.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %(root)r)
import torch
import __graft_entry__ as entry
pkg = entry.package()
text = open(os.path.join(%(root)r, "tests", "golden", "valeurs")).read()
st = torch.cuda.current_stream()
out13 = torch.zeros(13, dtype=torch.float64, device="cuda")
for features, sizes in (("", (1000000, 125000)), ("standard-random,f32", (1000000, 125000))):
    for n in sizes:
        sim = pkg.Simulator(pkg.Configuration.parse(text, features).with_num_events(n * 10000))
        sim.set_stream(st.cuda_stream)
        best = 1e30
        for rep in range(3):
            for _ in range(2): sim.simulate_merged_device(0, n, 10000, out13.data_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            K = 5 if n >= 1000000 else 20
            e0.record(st)
            for _ in range(K): sim.simulate_merged_device(0, n, 10000, out13.data_ptr())
            e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / K)
        print(f"  fused kernel {features!r:22} {n:8d} batches: {best:8.3f} ms  {n * 1e4 / best / 1e-3:.4g} events/s", flush=True)
        sim.close()
n_events = 2 * 10**9
cfg = pkg.Configuration.parse(text, "faster-evgen,no-photon-sorting").with_num_events(n_events)
nb, last = pkg.batch_layout(n_events)
with pkg.Simulator(cfg) as sim:
    sim.simulate_merged(0, 2000, 10000)
    best = 1e30
    for _ in range(4):
        t0 = time.perf_counter()
        acc = sim.simulate_merged(0, nb, last)
        best = min(best, time.perf_counter() - t0)
    print(f"  faster-evgen,no-photon-sorting 2e9 events: {best * 1e3:8.2f} ms  {n_events / best:.4g} events/s  selected {acc.selected_events}", flush=True)
'''
for lib in sys.argv[1:]:
    print(f"# {lib}", flush=True)
    env = dict(os.environ, TP3_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout + (r.stderr[-2000:] if r.returncode else ""), flush=True)
