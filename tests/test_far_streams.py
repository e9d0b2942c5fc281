"""Random streams at the batch indices the benchmark actually reaches (VERDICT r01 "stream-parity hole").

tests/golden/far_stream_kat.json holds known answers made by oracle/gen_far_kat.cpp, which steps the ORACLE's generators
sequentially, one random() at a time, from the seed to batches 999 999 / 1 968 526 / 1 968 527 / 7 999 999 (RANF; batch
1 968 527 is the first whose round index needs the 5th byte digit of the device jump-ahead, 2^32 rounds) and 999 999 /
7 999 999 (xoshiro256+ / xoshiro128+, sequential stream and jump() seeding) -- no jump-ahead algebra anywhere in the
generator.  Checked here:
  * CPU (`-m "not gpu"`): the host mirror of the device jump-ahead (tp3_host_ranf_round / tp3_host_xoshiro_state, same
    tables and algebra as rng.cuh) against those states; an INDEPENDENT pure-Python x^n mod P (square and multiply, no
    digit tables) against them too, and then -- pinned that way -- as the known answer for a batch near 3e8, the limit
    of the RANF tables, which no sequential walk reaches in reasonable time.
  * GPU (`-m gpu`): tp3_rng_dump (the simulation kernel's own stream code) equals the 240 words of every entry, and the
    batch accumulators at batches 999 999 / 7 999 999 equal the oracle's, fast-forwarded there, to 1e-10.
"""
import json
import os

import pytest

from conftest import ROOT

KAT_PATH = os.path.join(ROOT, "tests", "golden", "far_stream_kat.json")
MOD = 10**9
DRAWS_PER_BATCH = 120_000


def _kat():
    with open(KAT_PATH) as f:
        return json.load(f)["entries"]


def _features(e):
    f = []
    if e["rng"] != "ranf":
        f.append("standard-random")
    if e["dtype"] == "f32":
        f.append("f32")
    if e["seeding"] == "jump":
        f += ["multi-threading", "faster-threading"]
    return ",".join(f)


# ----------------------------------------------------------------------------- independent RANF jump (pure Python)
def _polmulmod(a, b):
    """a * b mod (x^55 + x^31 - 1) over Z/1e9: x^55 = 1 - x^31 (ranf.rs:106-119 as y[m+55] = y[m] - y[m+31])."""
    prod = [0] * 109
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                prod[i + j] = (prod[i + j] + ai * bj) % MOD
    for k in range(108, 54, -1):
        c = prod[k]
        if c:
            prod[k] = 0
            prod[k - 55] = (prod[k - 55] + c) % MOD
            prod[k - 24] = (prod[k - 24] - c) % MOD
    return prod[:55]


def _x_pow(n):
    result, base = [1] + [0] * 54, [0, 1] + [0] * 53
    while n:
        if n & 1:
            result = _polmulmod(result, base)
        base = _polmulmod(base, base)
        n >>= 1
    return result


def _seeded_round0(seed=234612947):
    """ranf.rs:36-66: IN55 initialisation + 10 warm-up rounds; returns numbers[1..55]."""
    n = [0] * 56
    n[55] = seed
    j, k = seed, 1
    for i in range(1, 55):
        ii = (21 * i) % 55
        n[ii] = k
        k = j - k
        if k < 0:
            k += MOD
        j = n[ii]
    for _ in range(10):
        _reset(n)
    return n[1:]


def _reset(n):
    for i in range(1, 25):
        n[i] = (n[i] - n[i + 31]) % MOD
    for i in range(25, 56):
        n[i] = (n[i] - n[i - 24]) % MOD


def python_ranf_round(rho):
    """numbers[1..55] after `rho` resets from the seeded state, by polynomial jump-ahead in exact integer arithmetic."""
    y = _seeded_round0()  # y[0..54] = y_1..y_55
    ext = y + [0] * 54
    for m in range(55, 109):  # y_{m+1} = y_{m+1-55} - y_{m+1-24}
        ext[m] = (ext[m - 55] - ext[m - 24]) % MOD
    c = _x_pow(55 * rho)
    return [sum(c[j] * ext[i + j] for j in range(55)) % MOD for i in range(55)]


def _expected_round_and_index(batch):
    d = batch * DRAWS_PER_BATCH
    rho, q = divmod(d, 55)
    # the reference resets lazily (ranf.rs:87-92): at a round boundary the generator still holds the consumed round
    return (rho - 1, 0) if q == 0 and d > 0 else (rho, 55 - q)


def _words_from_round(numbers, index, n_words, next_round):
    """The draws the reference hands out from (numbers[1..55], index): slots index, index-1, ... then the next round."""
    out, idx, cur = [], index, list(numbers)
    rho_next = next_round
    while len(out) < n_words:
        if idx == 0:
            cur, idx = python_ranf_round(rho_next), 55
            rho_next += 1
        out.append(cur[idx - 1])
        idx -= 1
    return out


# ----------------------------------------------------------------------------- CPU checks
def test_kat_file_is_complete():
    got = {(e["rng"], e["seeding"], e["batch"]) for e in _kat()}
    want = {("ranf", "sequential", b) for b in (999_999, 1_968_526, 1_968_527, 7_999_999)}
    for g in ("xoshiro256+", "xoshiro128+"):
        want |= {(g, "sequential", 999_999), (g, "sequential", 7_999_999), (g, "jump", 7_999_999)}
    assert want <= got


def test_host_and_python_jump_ahead_match_the_sequential_walk(tp3):
    import ctypes as C
    lib = tp3.lib()
    for e in _kat():
        if e["rng"] != "ranf":
            continue
        rho, index = _expected_round_and_index(e["batch"])
        assert e["state"][55] == index, e["batch"]
        out = (C.c_uint32 * 55)()
        assert lib.tp3_host_ranf_round(234612947, rho, out) == 0
        assert list(out) == e["state"][:55], f"host jump-ahead, batch {e['batch']}"
        assert python_ranf_round(rho) == e["state"][:55], f"python jump-ahead, batch {e['batch']}"
        assert _words_from_round(e["state"][:55], index, 240, rho + 1) == e["words"]


def test_host_xoshiro_jump_ahead_matches_the_sequential_walk(tp3):
    import ctypes as C
    lib = tp3.lib()
    for e in _kat():
        if e["rng"] == "ranf":
            continue
        out = (C.c_uint64 * 4)()
        steps, jumps = (e["batch"] * DRAWS_PER_BATCH, 0) if e["seeding"] == "sequential" else (0, e["batch"])
        assert lib.tp3_host_xoshiro_state(1 if e["dtype"] == "f32" else 0, steps, jumps, out) == 0
        assert list(out) == e["state"], (e["rng"], e["seeding"], e["batch"])


FAR_BATCH = 299_999_999  # the last batch the RANF jump tables reach (api.cu: 3e8 batches)


def test_host_jump_ahead_near_the_table_limit_matches_python(tp3):
    """No sequential walk gets to 3.6e13 draws; the independent Python jump-ahead (pinned on the sequential known answers
    above) is the known answer there."""
    import ctypes as C
    rho, _ = _expected_round_and_index(FAR_BATCH)
    out = (C.c_uint32 * 55)()
    assert tp3.lib().tp3_host_ranf_round(234612947, rho, out) == 0
    assert list(out) == python_ranf_round(rho)


# ----------------------------------------------------------------------------- GPU checks
@pytest.mark.gpu
def test_device_streams_match_the_sequential_walk(tp3, valeurs_text):
    sims = {}
    try:
        for e in _kat():
            f = _features(e)
            if f not in sims:
                sims[f] = tp3.Simulator(tp3.Configuration.parse(valeurs_text, f))
            assert sims[f].rng_dump(e["batch"], 240) == e["words"], (e["rng"], e["seeding"], e["batch"])
    finally:
        for s in sims.values():
            s.close()


@pytest.mark.gpu
def test_device_stream_near_the_table_limit(tp3, valeurs_text):
    rho, index = _expected_round_and_index(FAR_BATCH)
    want = _words_from_round(python_ranf_round(rho), index, 240, rho + 1)
    with tp3.Simulator(tp3.Configuration.parse(valeurs_text)) as sim:
        assert sim.rng_dump(FAR_BATCH, 240) == want


@pytest.mark.gpu
def test_device_accumulators_at_far_batches_match_the_oracle(tp3, valeurs_text):
    from test_gpu_parity import REL_F64, assert_acc_close
    sims = {}
    n_checked = 0
    try:
        for e in _kat():
            if "acc" not in e:
                continue
            f = _features(e)
            if f not in sims:
                cfg = tp3.Configuration.parse(valeurs_text, f).with_num_events(e["acc"]["num_events_total"])
                sims[f] = tp3.Simulator(cfg)
            got = sims[f].simulate_batches(e["batch"], 1)[0]
            want = tp3.Acc()
            want.selected_events = e["acc"]["selected_events"]
            for k in range(5):
                want.spm2[k], want.vars[k] = e["acc"]["spm2"][k], e["acc"]["vars"][k]
            want.sigma, want.variance = e["acc"]["sigma"], e["acc"]["variance"]
            assert got.selected_events == want.selected_events, (e["rng"], e["batch"])
            assert_acc_close(got, want, REL_F64, what=f"{e['rng']} batch {e['batch']}")
            n_checked += 1
    finally:
        for s in sims.values():
            s.close()
    assert n_checked >= 4
