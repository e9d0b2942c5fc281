#!/usr/bin/env python
"""Dynamic schedule with and without ramp units (device time, CUDA events). usage: ramp_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
pkg = entry.package()
text = open(os.path.join(ROOT, "tests", "golden", "valeurs")).read()
for n in (125000, 1000000):
    cfg = pkg.Configuration.parse(text).with_num_events(n * 10000)
    for ramp in (0, -1, 2368 * 2):
        with pkg.Simulator(cfg) as sim:
            st = torch.cuda.current_stream()
            sim.set_stream(st.cuda_stream)
            sim.set_option("ramp_units", ramp)
            out13 = torch.zeros(13, dtype=torch.float64, device="cuda")
            for _ in range(3):
                sim.simulate_merged_device(0, n, 10000, out13.data_ptr())
            torch.cuda.synchronize()
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                K = 8 if n < 500000 else 3
                for _ in range(K):
                    sim.simulate_merged_device(0, n, 10000, out13.data_ptr())
                e1.record(st)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / K)
            print(f"{n} batches, ramp_units {ramp:5d} (-1 = one wave): {best:8.3f} ms  {n * 1e4 / best / 1e-3:.4g} events/s", flush=True)
