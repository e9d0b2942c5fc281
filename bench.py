#!/usr/bin/env python
"""bench.py — events/s of the 3photons hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle restatement)

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): ONE run of the default
`valeurs` with num_events = 10^10, i.e. 10^6 batches of 10 000 events (scheduling/mod.rs:21), default
features (f64, RANF, sorted photons), its batch range sharded contiguously over the N ranks with no
data-path collective (multi_threading.rs:25,46-70): strong scaling, the default.  `--scaling weak` gives
every GPU 10^10 events instead (N x 10^10 in total) and is reported as a sub-record of the same line.

A "step" is one pass of the fused kernel (with its in-kernel ordered fold) over this rank's batch range.
  value : whole-job events/s, device time (CUDA events on the launching stream, max over ranks),
          everything the kernel reads (jump table, parameters) already resident in HBM; the merged
          accumulator stays in HBM.
  e2e   : the same metric through the call INTEGRATION.md tells a maintainer to bind: tp3_simulate_merged
          into a HOST accumulator (N = 1), or tp3_simulate_merged_device + ONE ncclReduce(sum) of the 13
          doubles to rank 0 + a 104-byte device->host copy (N > 1), then finalize() on the host.
  e2e_per_batch : the per-batch form of the boundary: one accumulator per batch into a host array AND the
          merged result, finalize()d -- value through tp3_simulate_batches_merged (the ordered fold comes from
          the same launch), host_fold_value through tp3_simulate_batches + the host left fold tp3_fold_batches.
  roofline: the path is FP64-pipe bound (no tensor cores, ~0.01 B/event of HBM traffic).  `achieved` /
          `frac` follow SURVEY.md section 8d: algorithmic FP64 TFLOP/s (833 flop per generated event as
          the reference writes them) over the DFMA peak measured live on this GPU by tp3_peak_probe
          (MEASURED_PEAKS.json carries no FP64 figure).  Because the fast kernel EXECUTES fewer flops
          than the reference writes, `executed_tflops` / `frac_executed` (SASS-counted FP64 flops per
          event from the committed ncu capture) and `pipe_active` (ncu FP64-pipe active fraction) are
          given next to it.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FLOP_PER_EVENT = 833.0  # 203 (generation + cuts) + 0.7082 * 890 (matrix elements)
REFERENCE_FEATURES = "multi-threading"  # the reference's reproducible multi-threaded build


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--events", type=float, default=1e10, help="events per step in total (strong) or PER GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--no-weak-subrecord", action="store_true", help="N > 1, strong: skip the extra weak-scaling measurement")
    ap.add_argument("--features", default="", help="cargo-style feature list, e.g. f32,standard-random")
    ap.add_argument("--kernel", default="fast", choices=["fast", "literal"])
    ap.add_argument("--cpu-sample-events", type=float, default=2e8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def valeurs_text():
    with open(os.path.join(ROOT, "tests", "golden", "valeurs")) as f:
        return f.read()


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


ORACLE_NATIVE_FLAGS = ["-O3", "-march=native", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-pthread"]
_oracle_build = None


def native_oracle():
    """The CPU baseline is rebuilt ON THIS BOX with -O3 -march=native (-ffp-contract=off stays: the
    Rust reference never fuses a*b+c).  Falls back to the portable prebuilt library if there is no compiler."""
    global _oracle_build
    if _oracle_build is None:
        import oracle_lib  # CPU baseline leg: the only place bench.py executes oracle/
        out = os.path.join(ROOT, "oracle", "_build", "liboracle_native.so")
        cmd = ["g++"] + ORACLE_NATIVE_FLAGS + ["-shared", "-o", out, os.path.join(ROOT, "oracle", "oracle_main.cpp")]
        try:
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.run(cmd, check=True, capture_output=True, timeout=300)
            oracle_lib.use_library(out)
            _oracle_build = "g++ " + " ".join(ORACLE_NATIVE_FLAGS) + " (built on this box)"
        except (OSError, subprocess.SubprocessError) as e:
            _oracle_build = f"prebuilt -O2 library (native rebuild failed: {type(e).__name__})"
    return _oracle_build


def cpu_reference_rate(n_events, threads):
    """events/s of the CPU oracle run with the reference's `multi-threading` semantics (one task
    per batch, scheduler thread pre-advances the RNG, ordered merge) on `threads` host threads."""
    import oracle_lib  # CPU baseline leg: the only place bench.py executes oracle/
    native_oracle()
    run = oracle_lib.run(valeurs_text(), REFERENCE_FEATURES, threads=threads, num_events=int(n_events),
                         want_batches=False, want_text=False)
    return n_events / run.seconds, run.seconds


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU with nvidia-smi during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < 3.0:  # until nvidia-smi delivers (outside the timed region)
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark(self):
        """The timed region starts now: only samples taken from here on count (nvidia-smi itself needs ~0.5 s to deliver
        its first sample, so the sampler is started before the warm-up steps)."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        rows = [r for t, r in self.rows if t >= getattr(self, "t_mark", 0.0) and len(r) >= 7 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i] == "Active" for r in rows)]
        power = max(float(r[2]) for r in rows if r[2].replace(".", "").isdigit())
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "power_w_max": power}


def config_dict(args, world, n_events, nb):
    """`config` of the JSON line: the same dictionary for both arms (the reference arm times a bounded sample of it)."""
    strong = args.scaling == "strong"
    workload = (f"default valeurs, ONE run of num_events={n_events:.3g} ({nb} batches of 10000; BASELINE configs[4], the shape of "
                f"configs[1]) sharded over {world} GPU(s) by contiguous batch ranges") if strong else (
                f"default valeurs, num_events={n_events:.3g} = {args.events:.3g} per GPU ({nb} batches of 10000), contiguous batch ranges per GPU")
    return {"workload": workload, "features": args.features or "default (f64, RANF, photon sorting)", "kernel": args.kernel,
            "l2": "not applicable: no input tensors; the kernel reads a 281 KB jump table and writes 104 B per batch"}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path, all host threads."""
    if rank != 0:
        return
    threads = host_threads()
    sample = args.cpu_sample_events
    n_total = int(args.events) * (args.gpus if args.scaling == "weak" else 1)
    for _ in range(args.warmup):
        cpu_reference_rate(min(sample, 2e6), threads)
    t = 0.0
    for _ in range(args.steps):
        _, s = cpu_reference_rate(sample, threads)
        t += s
    rate = sample * args.steps / t
    unit = "events/s"
    desc = (f"{sample:.3g} events of the default valeurs per step (bounded sample of the 1e10-event workload), "
            f"C++ restatement of the reference's `multi-threading` build (oracle/), {threads} threads, {native_oracle()}")
    print(json.dumps({
        "impl": "reference", "metric": "events/sec", "value": rate, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, args.gpus, n_total, (n_total + 9999) // 10000),
        "reference_build": f"cargo features `{REFERENCE_FEATURES}` (same numbers as the default features: the reference's golden is shared)",
        "cpu_baseline": {"value": rate, "unit": unit, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# FP64 work the shipped fast f64 kernel EXECUTES per generated event (SASS-counted from the committed ncu capture:
# DFMA = 2 flops, DMUL / DADD = 1), and the ncu FP64-pipe active fraction of the bench's own launch.
EXECUTED = {"source": "profiles/r02_bench_kernel_1e10_events.txt (ncu --set full of this launch: 1e6 batches, in-kernel fold on)",
            "dfma": 161.2, "dmul": 104.6, "dadd": 54.0, "pipe_active": 0.682, "traffic": 3760896 + 60777984}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    pkg = entry.package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    kernel = pkg.KERNEL_FAST if args.kernel == "fast" else pkg.KERNEL_LITERAL
    stream = torch.cuda.current_stream()
    warmup = max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def over_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def measure(scaling, steps, with_e2e):
        """One configuration: strong = ONE run of args.events events sharded over the ranks (configs[4] as written),
        weak = args.events per GPU."""
        n_events = int(args.events) * (world if scaling == "weak" else 1)
        cfg = pkg.Configuration.parse(valeurs_text(), args.features).with_num_events(n_events)
        nb, last = pkg.batch_layout(n_events)
        lo, cnt = pkg.shard_range(nb, world, rank)  # multi_threading.rs:25,46-70: contiguous batch ranges
        my_last = last if lo + cnt == nb else pkg.EVENT_BATCH_SIZE
        sim = pkg.Simulator(cfg, kernel, devices=[local_rank])
        sim.set_stream(stream.cuda_stream)
        out13 = torch.zeros(13, dtype=torch.float64, device="cuda")
        res = {"n_events": n_events, "nb": nb, "cfg": cfg, "sim": sim}

        # faster-evgen on the sequential RANF stream at N > 1: the STREAM is sharded, not the batches (tp3_fe_tile_device):
        # a rank cannot reach "its" batches without walking everything before them, but it can take a range of rounds
        f = set(args.features.split(","))
        fe_tiles = world > 1 and "faster-evgen" in f and "standard-random" not in f and not {"multi-threading", "faster-threading"} <= f

        def tile13(first_round, n_rounds, max_events):
            return out13, sim.fe_tile_device(first_round, n_rounds, max_events, out13.data_ptr())

        def step_device():
            if fe_tiles:  # includes the 8-byte count exchange and the reduce of 13 doubles: there is no device-only form
                return pkg.run_simulation_tiles(cfg, tile13, world, rank, dist, "cuda")
            sim.simulate_merged_device(lo, cnt, my_last, out13.data_ptr())

        # ---- device-resident timing ("value") ----
        sampler = ClockSampler(local_rank)
        sampler.start()
        for _ in range(warmup):
            step_device()
        barrier()
        launches0 = sim.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler.mark()
        ev0.record(stream)
        for _ in range(steps):
            step_device()
        ev1.record(stream)
        barrier()
        res["dev_ms"] = over_ranks(ev0.elapsed_time(ev1), dist.ReduceOp.MAX)
        res["launches"] = int(over_ranks(sim.launch_count - launches0, dist.ReduceOp.SUM))
        res["clocks"] = sampler.stop()
        res["value"] = n_events * steps / (res["dev_ms"] * 1e-3)
        if not with_e2e:
            return res

        # ---- end to end through the C ABI ("e2e"): what a caller of the reference's run_simulation gets — the merged
        # accumulator of the whole run on the HOST, finalized.
        def merged13(first, n_b, last_len):
            sim.simulate_merged_device(first, n_b, last_len, out13.data_ptr())
            return out13

        def e2e_step():
            if fe_tiles:
                fin_ = step_device()
                if rank != 0:
                    torch.cuda.synchronize()
                return fin_
            if world == 1:
                return pkg.finalize(cfg, sim.simulate_merged(lo, cnt, my_last))
            # tp3_simulate_merged_device + the run's one inter-GPU exchange (ncclReduce of 13 doubles) + 104-byte D2H + finalize
            fin_ = pkg.run_simulation_reduced(cfg, merged13, world, rank, dist)
            if rank != 0:
                torch.cuda.synchronize()
            return fin_

        for _ in range(2):
            fin = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fin = e2e_step()
        barrier()
        e2e_s = over_ranks(time.perf_counter() - t0, dist.ReduceOp.MAX)
        res["e2e"] = n_events * steps / e2e_s
        res["fin"] = fin

        # ---- the per-batch form of the boundary: one accumulator per batch into a host array + host left fold ----
        host = (pkg.Acc * cnt)()

        def per_batch_step(host_fold):
            if fe_tiles:  # per-batch accumulators by absolute batch index need the whole prefix of the stream: not sharded
                return e2e_step()
            if host_fold:  # the north star's wording: one accumulator per batch to the host, the host merges in the reference's order
                sim._check(pkg.lib().tp3_simulate_batches(sim._h, lo, cnt, my_last, host))
                mine = pkg.fold(host, cfg.flags)
            else:          # the same accumulators on the host, their ordered fold from the same launch (tp3_simulate_batches_merged)
                mine = pkg.Acc()
                sim._check(pkg.lib().tp3_simulate_batches_merged(sim._h, lo, cnt, my_last, host, ctypes.byref(mine)))
            if world == 1:
                return pkg.finalize(cfg, mine)
            t = torch.frombuffer(bytearray(bytes(mine)), dtype=torch.uint8).cuda()
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            if rank != 0:
                return None
            accs = [pkg.Acc.from_buffer_copy(p.cpu().numpy().tobytes()) for p in parts]
            return pkg.finalize(cfg, pkg.fold(accs, cfg.flags))

        n_pb = max(1, min(steps, 5))
        for host_fold, key in ((False, "e2e_per_batch"), (True, "e2e_per_batch_host_fold")):
            per_batch_step(host_fold)
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_pb):
                fin_pb = per_batch_step(host_fold)
            barrier()
            res[key] = n_events * n_pb / over_ranks(time.perf_counter() - t0, dist.ReduceOp.MAX)
        res["fin_pb"] = fin_pb
        res["d2h_per_batch"] = nb * ctypes.sizeof(pkg.Acc)
        return res

    main_res = measure(args.scaling, args.steps, True)
    sim = main_res["sim"]
    peak_tflops = sim.peak_probe(1 if "f32" in args.features else 0)
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak_subrecord:
        w = measure("weak", max(2, min(args.steps, 5)), False)
        weak = {"value": w["value"], "unit": "events/s", "events_per_gpu": int(args.events), "ms_per_step": w["dev_ms"] / max(2, min(args.steps, 5))}
        w["sim"].close()

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        cpu_reference_rate(2e6, threads)
        rate, secs = cpu_reference_rate(args.cpu_sample_events, threads)
        cpu = {"value": rate, "unit": "events/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_sample_events:.3g} events of the default valeurs ({secs:.2f} s), C++ restatement of "
                         f"the reference's `multi-threading` build, {threads} threads, {native_oracle()}"}

    if rank == 0:
        n_events, nb, cfg, value, dev_ms = main_res["n_events"], main_res["nb"], main_res["cfg"], main_res["value"], main_res["dev_ms"]
        achieved = value / world * FLOP_PER_EVENT / 1e12  # per GPU, to compare with a per-GPU peak
        f32 = "f32" in args.features
        default_f64 = not args.features and args.kernel == "fast"
        acc_bytes = ctypes.sizeof(pkg.Acc)
        exec_flop = 2 * EXECUTED["dfma"] + EXECUTED["dmul"] + EXECUTED["dadd"]
        # bytes per launch, from the committed ncu capture of exactly this launch shape (default features, 1e6 batches)
        ncu_traffic = EXECUTED["traffic"] if (default_f64 and nb // world == 1000000) else None
        fin, fin_pb = main_res["fin"], main_res["fin_pb"]
        line = {
            "metric": "events/sec", "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32" if f32 else "f64", "data": "synthetic",
            "config": config_dict(args, world, n_events, nb),
            "clocks": main_res["clocks"],
            "e2e": {"value": main_res["e2e"], "unit": "events/s", "h2d_bytes_per_step": world * int(pkg.lib().tp3_kernel_arg_bytes()),
                    "d2h_bytes_per_step": acc_bytes,
                    "note": "tp3_simulate_merged into a host tp3_acc + finalize() (N = 1); tp3_simulate_merged_device + one ncclReduce(sum) of 13 "
                            "doubles + 104-byte D2H on rank 0 + finalize() (N > 1). The only host->device payload is the kernel argument block."},
            "e2e_per_batch": {"value": main_res["e2e_per_batch"], "unit": "events/s", "d2h_bytes_per_step": main_res["d2h_per_batch"],
                              "host_fold_value": main_res["e2e_per_batch_host_fold"],
                              "note": "value: tp3_simulate_batches_merged = one 104-byte accumulator per batch into a host array (copied while the kernel "
                                      "runs) + their ordered fold from the same launch + finalize(); host_fold_value: tp3_simulate_batches into the "
                                      "host array + tp3_fold_batches on the host + finalize()"},
            "gpu_launches": main_res["launches"],
            "roofline": {"bound": "fp64" if not f32 else "fp32", "achieved": achieved, "peak": peak_tflops,
                         "unit": "TFLOP/s", "frac": achieved / peak_tflops, "traffic": ncu_traffic,
                         "frac_note": "SURVEY.md section 8d convention: 833 flop per generated event AS THE REFERENCE WRITES THEM over the measured DFMA peak; "
                                      "the fast kernel executes fewer (see executed_*), so this is reference-equivalent throughput, not pipe utilisation",
                         "executed_flop_per_event": exec_flop if default_f64 else None,
                         "executed_tflops": value / world * exec_flop / 1e12 if default_f64 else None,
                         "frac_executed": value / world * exec_flop / 1e12 / peak_tflops if default_f64 else None,
                         "pipe_active": EXECUTED["pipe_active"] if default_f64 else None,
                         "executed_source": EXECUTED["source"] if default_f64 else None,
                         "traffic_note": None if ncu_traffic is None else
                                         "DRAM bytes of one launch of the default f64 kernel over 1e6 batches, ncu --set full: the jump table read + "
                                         "the part of the 104 MB of per-batch accumulators that leaves L2; the bound is the FP64 pipe, not HBM",
                         "peak_source": "measured live: tp3_peak_probe (8 independent FMA chains per thread, all SMs)",
                         "flop_per_event": FLOP_PER_EVENT},
            "cpu_baseline": cpu,
            "weak": weak,
            "check": {"selected_events": fin.selected_events, "sigma_pb": fin.sigma,
                      "per_batch_path_selected_events": fin_pb.selected_events, "per_batch_path_sigma_pb": fin_pb.sigma},
        }
        print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
