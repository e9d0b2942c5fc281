#!/usr/bin/env python
"""What the in-kernel ordered fold costs with the SHIPPED schedule (device time, CUDA events, one B200): the same launch with
and without the fold, at the sizes one GPU gets of the 1e10-event run on 1 and on 8 GPUs.  usage: fold_cost_probe.py [features]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

pkg = entry.package()
features = sys.argv[1] if len(sys.argv) > 1 else ""
text = open(os.path.join(ROOT, "tests", "golden", "valeurs")).read()
st = torch.cuda.current_stream()
out13 = torch.zeros(13, dtype=torch.float64, device="cuda")
for n in (1000000, 125000):
    cfg = pkg.Configuration.parse(text, features).with_num_events(n * 10000)
    sim = pkg.Simulator(cfg)
    sim.set_stream(st.cuda_stream)
    res = {}
    for rep in range(3):
        for mode in ("nofold", "fold"):
            step = (lambda: sim.simulate_batches_device(0, n)) if mode == "nofold" else (lambda: sim.simulate_merged_device(0, n, 10000, out13.data_ptr()))
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            K = 5 if n >= 1000000 else 20
            for _ in range(K):
                step()
            e1.record(st)
            torch.cuda.synchronize()
            res.setdefault(mode, []).append(e0.elapsed_time(e1) / K)
    a, b = min(res["nofold"]), min(res["fold"])
    print(f"{n:8d} batches [{features!r}]: no fold {a:8.3f} ms   in-kernel fold {b:8.3f} ms   (+{100 * (b / a - 1):.2f} %)   all: {[round(x, 3) for x in res['nofold']]} {[round(x, 3) for x in res['fold']]}", flush=True)
