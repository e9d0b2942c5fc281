"""The N > 1 path on CPU: a world_size-2 gloo process group runs run_simulation_distributed with
the ORACLE standing in for the GPU (no GPU here); the sharding, gather and ordered fold are the code
under test.  The result must be bit-identical to the single-process fold and reproduce the golden
file for the same event count."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

from conftest import ROOT, load_package

N_EVENTS = 45_000  # 5 batches over 2 ranks (2 + 3), the last one short


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_range_simulator(tp3, valeurs_text, n_events):
    import ctypes as C
    import oracle_lib

    def simulate_range(first, n, last_len):
        run = oracle_lib.run(valeurs_text, "", num_events=n_events, want_text=False)
        out = (tp3.Acc * n)()
        for i in range(n):
            C.memmove(C.byref(out[i]), C.byref(run.per_batch[first + i]), C.sizeof(tp3.Acc))
        # the range geometry handed to the GPU path must describe exactly these batches
        nb, last = tp3.batch_layout(n_events)
        assert last_len == (last if first + n == nb else tp3.EVENT_BATCH_SIZE)
        return out

    return simulate_range


def _worker(rank, world, port, valeurs_text, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    tp3 = load_package()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cfg = tp3.Configuration.parse(valeurs_text).with_num_events(N_EVENTS)
    fin = tp3.run_simulation_distributed(cfg, _oracle_range_simulator(tp3, valeurs_text, N_EVENTS), world, rank, dist, "cpu")
    # per-event observable histograms are reduced the same way (synthetic per-rank content here)
    local = tp3.Histograms(4, [[rank + 1 + b for b in range(4)] for _ in range(tp3.HIST_OBSERVABLES)],
                           [[0.5 * (rank + 1) * (b + 1) for b in range(4)] for _ in range(tp3.HIST_OBSERVABLES)])
    total = tp3.reduce_histograms(local, world, rank, dist, "cpu")
    # the order-insensitive merge: every rank folds its own range, one reduce(sum) of 13 doubles (the path bench.py times at N > 1)
    import torch
    simulate_range = _oracle_range_simulator(tp3, valeurs_text, N_EVENTS)

    def merged13(first, n, last_len):
        return torch.tensor(tp3.acc_to_f64x13(tp3.fold(simulate_range(first, n, last_len))), dtype=torch.float64)

    red = tp3.run_simulation_reduced(cfg, merged13, world, rank, dist)
    if rank == 0:
        q.put((fin.selected_events, fin.sigma, fin.res_data(), total.counts, total.weights, red.selected_events, red.sigma, red.res_data()))
    else:
        assert fin is None and total is None and red is None
    dist.barrier()
    dist.destroy_process_group()


def _tiles_worker(rank, world, port, valeurs_text, n_events, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib
    tp3 = load_package()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cfg = tp3.Configuration.parse(valeurs_text, "faster-evgen").with_num_events(n_events)

    def tile13(first_round, n_rounds, max_events):  # the oracle stands in for tp3_fe_tile_device
        acc, n = oracle_lib.fe_tile(valeurs_text, "faster-evgen", first_round, n_rounds, max_events)
        return torch.tensor(tp3.acc_to_f64x13(acc), dtype=torch.float64), n

    fin = tp3.run_simulation_tiles(cfg, tile13, world, rank, dist, "cpu")
    if rank == 0:
        q.put((fin.selected_events, fin.sigma))
    else:
        assert fin is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_faster_evgen_stream_tiles_over_ranks(tp3, oracle, valeurs_text, world):
    """faster-evgen at N > 1 shards the STREAM: run_simulation_tiles (tile bounds, the 8-byte count exchange, the last
    rank's remainder, the reduce of 13 doubles) over a world_size-2 gloo group, with the oracle's tile walk standing in
    for the GPU.  Same events as the sequential run: same selected-event count, sigma equal up to the order of additions."""
    n_events = 123_456
    cfg = tp3.Configuration.parse(valeurs_text, "faster-evgen").with_num_events(n_events)
    whole, n = oracle.fe_tile(valeurs_text, "faster-evgen", 0, 0, n_events)
    assert n == n_events
    want = tp3.finalize(cfg, tp3.Acc.from_buffer_copy(bytes(whole)))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tiles_worker, args=(r, world, port, valeurs_text, n_events, q)) for r in range(world)]
    for p in procs:
        p.start()
    sel, sigma = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sel == want.selected_events
    assert abs(sigma - want.sigma) <= 1e-12 * abs(want.sigma)


@pytest.mark.parametrize("world", [2])
def test_sharded_run_is_bit_identical_to_single_process(tp3, oracle, valeurs_text, world):
    cfg = tp3.Configuration.parse(valeurs_text).with_num_events(N_EVENTS)
    single = tp3.run_simulation_distributed(cfg, _oracle_range_simulator(tp3, valeurs_text, N_EVENTS), 1, 0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, valeurs_text, q)) for r in range(world)]
    for p in procs:
        p.start()
    sel, sigma, res, hist_counts, hist_weights, red_sel, red_sigma, red_res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (sel, sigma, res) == (single.selected_events, single.sigma, single.res_data())
    ranks = world * (world + 1) // 2  # sum of (rank + 1)
    assert hist_counts == [[ranks + world * b for b in range(4)]] * tp3.HIST_OBSERVABLES
    assert hist_weights == [[0.5 * ranks * (b + 1) for b in range(4)]] * tp3.HIST_OBSERVABLES
    # the reduced (fold of per-rank folds) result: same events, sums equal up to the association of one addition per field
    from numdiff import compare
    assert red_sel == sel and abs(red_sigma - sigma) <= 1e-14 * abs(sigma)
    assert compare(red_res, res, rel=1e-13) == []
    # and the fold agrees with the oracle's own whole-run text
    assert compare(res, oracle.run(valeurs_text, "", num_events=N_EVENTS).res_data) == []
