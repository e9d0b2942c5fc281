#!/usr/bin/env python
"""Device time of the default fused kernel (in-kernel fold) as a function of the launch size around the 8-GPU share of the
1e10-event run: T(n) against the straight line through the 1e6-batch rate shows what the launch shape costs at each n.
usage: tail_probe.py [features] [option=value ...] [--short | n=size,size,...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

pkg = entry.package()
features = sys.argv[1] if len(sys.argv) > 1 else ""
short = "--short" in sys.argv
sizes = [int(float(x)) for a in sys.argv[2:] if a.startswith("n=") for x in a[2:].split(",")]
opts = [a.split("=") for a in sys.argv[2:] if a != "--short" and not a.startswith("n=")]
text = open(os.path.join(ROOT, "tests", "golden", "valeurs")).read()
st = torch.cuda.current_stream()
out13 = torch.zeros(13, dtype=torch.float64, device="cuda")


def timed(n, K):
    sim = pkg.Simulator(pkg.Configuration.parse(text, features).with_num_events(n * 10000))
    sim.set_stream(st.cuda_stream)
    for k, v in opts:
        sim.set_option(k, int(v))
    best = 1e30
    for rep in range(2):
        for _ in range(2):
            sim.simulate_merged_device(0, n, 10000, out13.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(K):
            sim.simulate_merged_device(0, n, 10000, out13.data_ptr())
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K)
    sim.close()
    return best


ref = timed(1000000, 4) / 1000000
print(f"# features {features!r} options {opts}: {ref * 1e3:.4f} us per batch at 1e6 batches", flush=True)
for n in sizes if sizes else ([120768, 123136, 125000, 125504, 250000, 500000] if short else list(range(104192, 142081, 2368)) + [125000, 250000, 500000]):
    t = timed(n, 10)
    print(f"n = {n:7d}  T = {t:8.3f} ms   T - n * rate = {t - n * ref:6.3f} ms   efficiency {n * ref / t:.4f}", flush=True)
