"""CPU tests of the product's host side (no GPU, no compute): the C-ABI library loads and exports
every symbol of include/tp3.h, `valeurs` parsing follows config.rs, the kernel constants match the
oracle's, and finalize + res.data/stdout formatting reproduce the goldens when fed the oracle's
merged sums (the GPU tests feed them the GPU's sums)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, golden
from numdiff import compare


def test_library_exports_every_declared_symbol(tp3):
    header = open(os.path.join(ROOT, "include", "tp3.h")).read()
    declared = set(re.findall(r"\b(tp3_[a-z0-9_]+)\s*\(", header))
    assert declared == set(tp3.ABI_SYMBOLS)
    lib = tp3.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.tp3_abi_version() == 2


def test_struct_layouts_match_header(tp3):
    assert C.sizeof(tp3.Acc) == 104  # 13 scalars (resacc.rs:19-34)
    assert C.sizeof(tp3.Params) == 8 + 13 * 8 + 8
    assert C.sizeof(tp3.Config) == 8 + 14 * 8 + 3 * 4 + 4
    assert C.sizeof(tp3.Final) == 8 + 20 * 8 + 8 * 8


def test_no_gpu_fails_loudly(tp3, valeurs_text):
    """No CPU fallback: without a B200 the context cannot be created."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = tp3.Configuration.parse(valeurs_text)
    with pytest.raises(tp3.Tp3Error) as e:
        tp3.Simulator(cfg)
    assert e.value.code == tp3.E_NO_DEVICE


def test_config_parse_default(tp3, valeurs_text):
    cfg = tp3.Configuration.parse(valeurs_text)
    r = cfg.raw
    assert (r.num_events, r.e_total, r.beam_photons_cut, r.photon_photon_cut, r.e_min, r.beam_photon_plane_cut) == (
        10_000_000, 91.187, 0.9, 0.9396, 4.559, 0.0)
    assert (r.alpha, r.alpha_z, r.gev2_to_picobarn, r.m_z0, r.g_z0) == (7.297353079644818e-3, 7.8125e-3, 0.38937966e9, 91.187, 2.49)
    assert (r.sin2_weinberg, r.branching_ep_em, r.beta_plus, r.beta_minus, r.num_bins, r.impr, r.plot) == (
        0.2319, 0.03367, 1.0, 1.0, 200, 0, 0)


def test_config_parse_f32_rounds_text_directly(tp3, valeurs_text):
    import numpy as np
    cfg = tp3.Configuration.parse(valeurs_text, "f32")
    assert cfg.raw.alpha == float(np.float32("7.297353079644818e-3"))
    assert cfg.raw.gev2_to_picobarn == float(np.float32("0.38937966e9"))


@pytest.mark.parametrize("mutate,message", [
    (lambda t: "\n".join(t.splitlines()[:10]), "missing configuration of g_z0"),
    (lambda t: t.replace("10000000 ", "0 ", 1), "Please simulate at least one event"),
    (lambda t: t.replace("91.187e0", "abc", 1), "could not parse configuration of e_total"),
    (lambda t: re.sub(r"\.false\.(\s+'HBook)", r".true.\1", t), "Plotting is not supported"),
    (lambda t: re.sub(r"\.false\.(\s+'Impression)", r".TRUE.\1", t), "Individual result printing is not supported"),
    (lambda t: re.sub(r"\.false\.(\s+'Impression)", r"maybe\1", t), "could not parse configuration of impr"),
    # config.rs:187-195: only the FORTRAN forms are lower-cased; Rust's bool parser takes exactly "true" / "false"
    (lambda t: re.sub(r"\.false\.(\s+'Impression)", r"False\1", t), "could not parse configuration of impr"),
    (lambda t: re.sub(r"\.false\.(\s+'HBook)", r"TRUE\1", t), "could not parse configuration of plot"),
    # what strtod / strtol would take and Rust's FromStr does not
    (lambda t: t.replace("91.187e0", "0x1.6cbf7ced916873p+6", 1), "could not parse configuration of e_total"),
    (lambda t: t.replace("91.187e0", "nan(123)", 1), "could not parse configuration of e_total"),
    (lambda t: t.replace("10000000 ", "99999999999999999999999 ", 1), "could not parse configuration of num_events"),
    (lambda t: t.replace("10000000 ", "-5 ", 1), "could not parse configuration of num_events"),
    (lambda t: re.sub(r"^200(\s)", r"4294967296\1", t, flags=re.M), "could not parse configuration of num_bins"),
])
def test_config_errors(tp3, valeurs_text, mutate, message):
    with pytest.raises(tp3.Tp3Error) as e:
        tp3.Configuration.parse(mutate(valeurs_text))
    assert e.value.code == tp3.E_CONFIG and message in str(e.value)


def test_config_blank_lines_and_first_token_rule(tp3, valeurs_text):
    """config.rs:69-71: only the first whitespace-delimited token of each non-blank line counts."""
    lines = valeurs_text.splitlines()
    shuffled = "\n\n   \n".join("   " + l for l in lines[:18]) + "\n\n"
    a, b = tp3.Configuration.parse(valeurs_text).raw, tp3.Configuration.parse(shuffled).raw
    assert bytes(a) == bytes(b)


@pytest.mark.parametrize("features", ["", "f32"])
def test_kernel_constants_match_oracle(tp3, oracle, valeurs_text, features):
    p = tp3.Configuration.parse(valeurs_text, features).params()
    c = oracle.constants(valeurs_text, features)
    assert (p.g_a, p.g_beta_p, p.g_beta_m) == tuple(c[0:3])
    assert list(p.sigma_contribs) == c[5:10]
    assert p.num_events_total == 10_000_000


@pytest.mark.parametrize("features,suffix,tol", [
    ("", "", {}),
    ("f32", "f32", dict(abs_=1.1e-8)),
    ("standard-random", "standard-random", {}),
    ("multi-threading,faster-threading", "multi-threading,faster-threading", {}),
])
def test_finalize_and_text_from_oracle_sums(tp3, oracle, valeurs_text, features, suffix, tol):
    """Product finalize/formatters (host.cpp) on the oracle's per-batch sums == golden files."""
    run = oracle.run(valeurs_text, features, threads=8, want_text=False)
    cfg = tp3.Configuration.parse(valeurs_text, features)
    accs = (tp3.Acc * len(run.per_batch))()
    for i, b in enumerate(run.per_batch):
        C.memmove(C.byref(accs[i]), C.byref(b), C.sizeof(tp3.Acc))
    fin = tp3.finalize(cfg, tp3.fold(accs, cfg.flags))
    assert fin.selected_events == run.merged.selected_events
    assert compare(fin.res_data(), golden("res.data-features_" + suffix), **tol) == []
    out_tol = dict(rel=1.9e-5) if "f32" in features else {}
    assert compare(fin.stdout(), golden("stdout.log-features_" + suffix), **out_tol) == []


@pytest.mark.parametrize("features", ["", "f32"])
def test_finalize_and_text_randomised_against_oracle(tp3, oracle, valeurs_text, features):
    """The two independent restatements of finalize + the Rust-compatible number formatting (host.cpp, the product;
    oracle_text.hpp, the checker) must print the same text for sums the golden runs never produce: every decade,
    both signs, exact powers of ten, values that round up into the next decade, zero."""
    import random
    rnd = random.Random(20240607)
    cfg = tp3.Configuration.parse(valeurs_text, features)
    n = cfg.num_events

    def value(scale_pow):
        kind = rnd.random()
        if kind < 0.1:
            return 0.0
        if kind < 0.2:
            return rnd.choice([-1.0, 1.0]) * 10.0 ** rnd.randint(-12, 12)                    # exact powers of ten
        if kind < 0.3:
            return rnd.choice([-1.0, 1.0]) * (10.0 ** rnd.randint(-8, 8)) * (1 - 1e-15)      # rounds up into the next decade
        return rnd.choice([-1.0, 1.0]) * rnd.uniform(1.0, 10.0) * 10.0 ** (scale_pow + rnd.randint(-9, 9))

    for case in range(400):
        acc = tp3.Acc()
        acc.selected_events = rnd.randint(1, n)
        for k in range(5):
            acc.spm2[k] = value(3) if k >= 3 else abs(value(3))
            acc.vars[k] = abs(value(6)) + acc.spm2[k] ** 2 / n   # keeps the variances non-negative, as real sums are
        acc.sigma = abs(value(0)) + 1e-300
        acc.variance = abs(value(-6)) + acc.sigma ** 2 / n
        if "f32" in features:  # the sums of an f32 run are f32 values
            import struct
            for name in ("sigma", "variance"):
                setattr(acc, name, struct.unpack("f", struct.pack("f", getattr(acc, name)))[0])
            for k in range(5):
                acc.spm2[k] = struct.unpack("f", struct.pack("f", acc.spm2[k]))[0]
                acc.vars[k] = struct.unpack("f", struct.pack("f", min(acc.vars[k], 3e38)))[0]
        fin = tp3.finalize(cfg, acc)
        want_rd, want_so = oracle.finalize_text(valeurs_text, features, acc)
        assert fin.res_data() == want_rd, f"case {case}"
        assert fin.stdout() == want_so, f"case {case}"


def test_batch_layout_and_sharding(tp3):
    assert tp3.batch_layout(10_000_000) == (1000, 10000)
    assert tp3.batch_layout(10_001) == (2, 1)
    assert tp3.batch_layout(1) == (1, 1)
    for n, w in [(1000, 8), (7, 4), (1, 2), (10**6, 3)]:
        parts = [tp3.shard_range(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (lo, cnt), (lo2, _) in zip(parts, parts[1:]):
            assert lo + cnt == lo2


def test_host_ranf_jump_matches_oracle_stream(tp3, oracle):
    """Round r of the seeded generator via polynomial jump-ahead == the oracle stepped there.
    The reference hands out slots 55..1 of each round (ranf.rs:95-101)."""
    words = oracle.rng_words("", 0, 55 * 40)
    out = (C.c_uint32 * 55)()
    for r in (0, 1, 2, 39):
        assert tp3.lib().tp3_host_ranf_round(234612947, r, out) == 0
        assert list(out)[::-1] == words[55 * r: 55 * (r + 1)]
    # far jump: batch 999 starts at draw 119 880 000 = round 2 179 636, offset 20
    assert tp3.lib().tp3_host_ranf_round(234612947, 119_880_000 // 55, out) == 0
    off = 119_880_000 % 55
    assert list(out)[::-1][off:off + 4] == oracle.rng_words("", 999, 4)


@pytest.mark.parametrize("f32,features", [(0, "standard-random"), (1, "standard-random,f32")])
def test_host_xoshiro_jump_matches_oracle_stream(tp3, oracle, f32, features):
    st = (C.c_uint64 * 4)()
    mask = (1 << 32) - 1 if f32 else (1 << 64) - 1
    for batch in (0, 1, 7):
        assert tp3.lib().tp3_host_xoshiro_state(f32, 120_000 * batch, 0, st) == 0
        assert (st[0] + st[3]) & mask == oracle.rng_words(features, batch, 1)[0]
        assert tp3.lib().tp3_host_xoshiro_state(f32, 0, batch, st) == 0
        assert (st[0] + st[3]) & mask == oracle.rng_words(features + ",multi-threading,faster-threading", batch, 1)[0]


def test_missing_library_fails_loudly(tmp_path):
    """The product never falls back to anything: without the built CUDA library the package refuses to work."""
    import subprocess
    import sys
    code = ("import importlib.util, sys; spec = importlib.util.spec_from_file_location('tp3x', r'%s'); "
            "m = importlib.util.module_from_spec(spec); sys.modules['tp3x'] = m; spec.loader.exec_module(m); m.lib()"
            % os.path.join(ROOT, "3photons-rust_b200", "__init__.py"))
    env = dict(os.environ, TP3_LIB=str(tmp_path / "does_not_exist.so"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_config_accepts_what_rust_accepts(tp3, valeurs_text):
    """Rust's `bool` parser takes lower-case true/false; FORTRAN forms in any case; numbers with a leading +."""
    t = re.sub(r"\.false\.(\s+'Impression)", r"false\1", valeurs_text)
    t = re.sub(r"\.false\.(\s+'HBook)", r".FALSE.\1", t)
    t = t.replace("91.187e0", "+91.187", 1)
    a, b = tp3.Configuration.parse(valeurs_text).raw, tp3.Configuration.parse(t).raw
    assert bytes(a) == bytes(b)


def test_fold_batches_is_the_left_fold(tp3):
    """tp3_fold_batches = sequential.rs:24-36: starts FROM the first accumulator, adds the rest in order, in the run's Float."""
    import numpy as np
    rng = np.random.default_rng(7)
    n = 1000
    accs = (tp3.Acc * n)()
    vals = rng.standard_normal((n, 12)) * 10.0 ** rng.integers(-8, 8, (n, 12))
    for i in range(n):
        accs[i].selected_events = int(rng.integers(0, 10000))
        for k in range(5):
            accs[i].spm2[k], accs[i].vars[k] = vals[i, k], vals[i, 5 + k]
        accs[i].sigma, accs[i].variance = vals[i, 10], vals[i, 11]
    got = tp3.fold(accs)
    want = np.add.accumulate(vals, axis=0)[-1]
    assert list(got.spm2) + list(got.vars) + [got.sigma, got.variance] == want.tolist()
    assert got.selected_events == sum(a.selected_events for a in accs)
    got32 = tp3.fold(accs, tp3.F32)
    want32 = np.add.accumulate(vals.astype(np.float32), axis=0, dtype=np.float32)[-1]
    assert [np.float32(x) for x in list(got32.spm2) + list(got32.vars) + [got32.sigma, got32.variance]] == want32.tolist()
    one = tp3.fold(accs[:1])
    assert bytes(one) == bytes(accs[0])


def test_rust_sys_crate_matches_header(tp3):
    """rust/tp3-sys/src/lib.rs cannot be compiled here (no rustc in the image), so at least its extern "C" block is held
    to include/tp3.h statically: same symbol names, same number of arguments."""
    header = open(os.path.join(ROOT, "include", "tp3.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    decl = {m.group(1): m.group(2) for m in re.finditer(r"\b(tp3_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", header)}
    arity_h = {k: 0 if v.strip() in ("", "void") else v.count(",") + 1 for k, v in decl.items()}
    rust = open(os.path.join(ROOT, "rust", "tp3-sys", "src", "lib.rs")).read()
    rust = re.sub(r"//[^\n]*", "", rust)
    fns = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (tp3_[a-z0-9_]+)\s*\(([^)]*)\)", rust, flags=re.S)}
    arity_r = {k: 0 if not v.strip() else len([a for a in v.split(",") if a.strip()]) for k, v in fns.items()}
    assert set(arity_r) == set(arity_h), (set(arity_h) - set(arity_r), set(arity_r) - set(arity_h))
    assert arity_r == arity_h
    assert "pub const TP3_ABI_VERSION: i32 = %d;" % tp3.lib().tp3_abi_version() in rust


def test_merge_is_done_in_the_runs_float(tp3):
    """ResultsAccumulator::merge adds in Float (resacc.rs:133-139): under f32 the sums round to f32."""
    import numpy as np
    a, b = tp3.Acc(), tp3.Acc()
    a.selected_events, b.selected_events = 3, 4
    a.sigma, b.sigma = float(np.float32(1.0)), float(np.float32(1e-8))
    a.spm2[0], b.spm2[0] = 0.1, 0.2
    m64 = tp3.merge(tp3.Acc.from_buffer_copy(a), b, 0)
    assert m64.selected_events == 7 and m64.sigma == 1.0 + float(np.float32(1e-8)) and m64.spm2[0] == 0.1 + 0.2
    m32 = tp3.merge(tp3.Acc.from_buffer_copy(a), b, tp3.F32)
    assert m32.sigma == 1.0  # 1e-8 is below half an ulp of 1.0f
    assert m32.spm2[0] == float(np.float32(0.1) + np.float32(0.2))


@pytest.mark.parametrize("value,f64_text,f32_text", [
    (11.303932414679, "11.303932414679", "11.304"),
    (0.0028014060468836, "0.0028014060468836", "0.0028014"),
    (2.4782579584832e-4, "2.4782579584832e-4", "2.4783e-4"),
    (389379660.0, "389379660", "3.8938e8"),
    (0.0, "0", "0"),
    (128.0, "128", "128"),
    (-4.559, "-4.559", "-4.559"),
])
def test_engineering_format_matches_reference_goldens(tp3, valeurs_text, value, f64_text, f32_text):
    """output.rs:230-269 (`%g`-like writer) through the res.data text: put the value in the sigma slot."""
    for features, want in (("", f64_text), ("f32", f32_text)):
        cfg = tp3.Configuration.parse(valeurs_text, features)
        fin = tp3.Final()
        fin.sigma = value
        fin.prec = 1.0
        text = tp3.FinalResults(fin, cfg).res_data()
        line = [l for l in text.splitlines() if l.startswith(" Section Efficace")][0]
        assert line.split(":")[1].strip() == want


def test_every_option_is_documented_in_the_header():
    """Every name tp3_set_option / tp3_get_stat accepts (api.cu) is described in include/tp3.h."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "3photons-rust_b200", "csrc", "api.cu")).read()
    header = open(os.path.join(root, "include", "tp3.h")).read()
    names = sorted(set(re.findall(r'k == "([a-z_0-9]+)"', src)))
    assert len(names) >= 15, names
    missing = [n for n in names if f'"{n}"' not in header]
    assert not missing, missing
