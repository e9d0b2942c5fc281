"""3photons-rust_b200 — host-side mirror of the reference's run surface over libtp3.so.

The names follow the reference crate (`/root/reference/src`):

    Configuration.load      config.rs:57-128
    Couplings / params      coupling.rs:23-33, evgen.rs:40-77, resacc.rs:59-117
    ResultsAccumulator      resacc.rs:16-139 (one per 10 000-event batch)
    run_simulation          scheduling/mod.rs:31-59 (+ sequential.rs / multi_threading.rs semantics)
    FinalResults, dump_results   resfin.rs:26-194, output.rs:30-177

All compute goes through the C ABI of `include/tp3.h` (ctypes); there is no Python or CPU
implementation of the hot path in this package.  If the CUDA library has not been built, or no
B200 is visible, the calls fail loudly.

The package directory name is not a Python identifier; load it with
`importlib` (see `tests/conftest.py::load_package`) or via `__graft_entry__.package()`.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TP3_LIB") or os.path.join(_HERE, "_build", "libtp3.so")  # TP3_LIB: A/B builds during development
CLI_PATH = os.path.join(_HERE, "_build", "trois_photons_b200")

EVENT_BATCH_SIZE = 10_000  # scheduling/mod.rs:21

# Cargo features (Cargo.toml:9-21) as flag bits of include/tp3.h
F32 = 1 << 0
FASTER_EVGEN = 1 << 1
FASTER_THREADING = 1 << 2
MULTI_THREADING = 1 << 3
NO_PHOTON_SORTING = 1 << 4
STANDARD_RANDOM = 1 << 5
FEATURE_BITS = {
    "f32": F32,
    "faster-evgen": FASTER_EVGEN,
    "faster-threading": FASTER_THREADING,
    "multi-threading": MULTI_THREADING,
    "no-photon-sorting": NO_PHOTON_SORTING,
    "standard-random": STANDARD_RANDOM,
}
KERNEL_FAST, KERNEL_LITERAL = 0, 1
OK, E_INVALID, E_NO_DEVICE, E_CUDA, E_CONFIG, E_IO = range(6)


def feature_mask(features: "str | Iterable[str] | int") -> int:
    """'f32,standard-random' / ['f32', ...] / int -> flag bits. faster-threading is only
    meaningful together with multi-threading (scheduling/mod.rs:45-54), as in the reference."""
    if isinstance(features, int):
        mask = features
    else:
        if isinstance(features, str):
            features = [f for f in features.split(",") if f]
        mask = 0
        for f in features:
            if f not in FEATURE_BITS:
                raise ValueError(f"unknown feature {f!r}")
            mask |= FEATURE_BITS[f]
    if not mask & MULTI_THREADING:
        mask &= ~FASTER_THREADING
    return mask


class Tp3Error(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"tp3 error {code}: {message}")
        self.code = code


class Params(C.Structure):  # tp3_params
    _fields_ = [
        ("num_events_total", C.c_uint64),
        ("e_total", C.c_double),
        ("acut", C.c_double),
        ("bcut", C.c_double),
        ("e_min", C.c_double),
        ("sincut", C.c_double),
        ("g_a", C.c_double),
        ("g_beta_p", C.c_double),
        ("g_beta_m", C.c_double),
        ("sigma_contribs", C.c_double * 5),
        ("flags", C.c_uint32),
        ("kernel", C.c_uint32),
    ]


class Acc(C.Structure):  # tp3_acc == ResultsAccumulator's 13 sums (resacc.rs:19-34)
    _fields_ = [
        ("selected_events", C.c_uint64),
        ("spm2", C.c_double * 5),
        ("vars", C.c_double * 5),
        ("sigma", C.c_double),
        ("variance", C.c_double),
    ]

    def as_tuple(self):
        return (int(self.selected_events), tuple(self.spm2), tuple(self.vars), float(self.sigma), float(self.variance))


class Config(C.Structure):  # tp3_config == Configuration (config.rs:8-53)
    _fields_ = [
        ("num_events", C.c_uint64),
        ("e_total", C.c_double),
        ("beam_photons_cut", C.c_double),
        ("photon_photon_cut", C.c_double),
        ("e_min", C.c_double),
        ("beam_photon_plane_cut", C.c_double),
        ("alpha", C.c_double),
        ("alpha_z", C.c_double),
        ("gev2_to_picobarn", C.c_double),
        ("m_z0", C.c_double),
        ("g_z0", C.c_double),
        ("sin2_weinberg", C.c_double),
        ("branching_ep_em", C.c_double),
        ("beta_plus", C.c_double),
        ("beta_minus", C.c_double),
        ("num_bins", C.c_int32),
        ("impr", C.c_int32),
        ("plot", C.c_int32),
    ]


class Final(C.Structure):  # tp3_final == FinalResults (resfin.rs:26-62)
    _fields_ = [
        ("selected_events", C.c_uint64),
        ("spm2", (C.c_double * 5) * 2),
        ("vars", (C.c_double * 5) * 2),
        ("sigma", C.c_double),
        ("prec", C.c_double),
        ("variance", C.c_double),
        ("beta_min", C.c_double),
        ("ss_p", C.c_double),
        ("inc_ss_p", C.c_double),
        ("ss_m", C.c_double),
        ("inc_ss_m", C.c_double),
    ]


_lib = None


def lib() -> C.CDLL:
    """The C-ABI library. Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the hot path)"
        )
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    vp, u64, u32, i32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_double
    sigs = {
        "tp3_abi_version": (C.c_int, []),
        "tp3_create": (C.c_int, [P(Params), C.c_int, P(C.c_int), P(vp)]),
        "tp3_destroy": (None, [vp]),
        "tp3_last_error": (C.c_char_p, [vp]),
        "tp3_set_stream": (C.c_int, [vp, C.c_int, vp]),
        "tp3_simulate_batches": (C.c_int, [vp, u64, u64, u32, P(Acc)]),
        "tp3_simulate_batches_device": (C.c_int, [vp, u64, u64, u32]),
        "tp3_fetch": (C.c_int, [vp, P(Acc), u64]),
        "tp3_simulate_merged": (C.c_int, [vp, u64, u64, u32, P(Acc)]),
        "tp3_simulate_merged_device": (C.c_int, [vp, u64, u64, u32, vp]),
        "tp3_simulate_batches_merged": (C.c_int, [vp, u64, u64, u32, P(Acc), P(Acc)]),
        "tp3_fold_batches": (C.c_int, [P(Acc), u64, u32, P(Acc)]),
        "tp3_fe_tile_device": (C.c_int, [vp, u64, u64, u64, vp, P(u64)]),
        "tp3_set_option": (C.c_int, [vp, C.c_char_p, C.c_int64]),
        "tp3_get_stat": (C.c_int, [vp, C.c_char_p, P(C.c_int64)]),
        "tp3_kernel_arg_bytes": (C.c_size_t, []),
        "tp3_synchronize": (C.c_int, [vp]),
        "tp3_launch_count": (u64, [vp]),
        "tp3_histograms_enable": (C.c_int, [vp, u32]),
        "tp3_histograms_reset": (C.c_int, [vp]),
        "tp3_histograms_fetch": (C.c_int, [vp, P(u64), P(C.c_double)]),
        "tp3_rng_dump": (C.c_int, [vp, u64, u32, P(u64)]),
        "tp3_events_dump": (C.c_int, [vp, u64, u32, P(dbl), P(i32), P(dbl)]),
        "tp3_peak_probe": (C.c_int, [vp, C.c_int, P(dbl)]),
        "tp3_fastmath_probe": (C.c_int, [vp, C.c_int, u32, P(dbl), P(dbl)]),
        "tp3_config_parse": (C.c_int, [C.c_char_p, u32, P(Config), C.c_char_p, C.c_size_t]),
        "tp3_params_from_config": (C.c_int, [P(Config), u32, u32, P(Params)]),
        "tp3_merge": (C.c_int, [P(Acc), P(Acc), u32]),
        "tp3_finalize": (C.c_int, [P(Config), u32, P(Acc), P(Final)]),
        "tp3_format_res_data": (C.c_size_t, [P(Config), u32, P(Final), C.c_char_p, C.c_size_t]),
        "tp3_format_stdout": (C.c_size_t, [P(Config), u32, P(Final), C.c_char_p, C.c_size_t]),
        "tp3_run": (C.c_int, [C.c_char_p, C.c_char_p, u32, u32, C.c_int, C.c_char_p, C.c_size_t, P(dbl)]),
        "tp3_run_stages": (C.c_int, [C.c_char_p, C.c_char_p, u32, u32, C.c_int, C.c_char_p, C.c_size_t, P(dbl), P(dbl)]),
        "tp3_host_ranf_round": (C.c_int, [i32, u64, P(u32)]),
        "tp3_host_xoshiro_state": (C.c_int, [C.c_int, u64, u64, P(u64)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


ABI_SYMBOLS = [
    "tp3_abi_version", "tp3_create", "tp3_destroy", "tp3_last_error", "tp3_set_stream", "tp3_simulate_batches",
    "tp3_simulate_batches_device", "tp3_fetch", "tp3_simulate_merged", "tp3_synchronize", "tp3_launch_count",
    "tp3_rng_dump", "tp3_events_dump", "tp3_peak_probe", "tp3_fastmath_probe", "tp3_config_parse", "tp3_params_from_config", "tp3_merge",
    "tp3_finalize", "tp3_format_res_data", "tp3_format_stdout", "tp3_run", "tp3_host_ranf_round",
    "tp3_host_xoshiro_state", "tp3_histograms_enable", "tp3_histograms_reset", "tp3_histograms_fetch",
    "tp3_simulate_merged_device", "tp3_simulate_batches_merged", "tp3_fold_batches", "tp3_set_option", "tp3_get_stat", "tp3_kernel_arg_bytes", "tp3_run_stages",
    "tp3_fe_tile_device",
]


# --------------------------------------------------------------------------- host surface
class Configuration:
    """config.rs:8-128. `Configuration.load(path)` parses a `valeurs` file; `.echo` is what the
    reference prints at load time."""

    def __init__(self, raw: Config, flags: int):
        self.raw = raw
        self.flags = flags

    @classmethod
    def parse(cls, text: str, features="") -> "Configuration":
        flags = feature_mask(features)
        raw = Config()
        err = C.create_string_buffer(1024)
        rc = lib().tp3_config_parse(text.encode(), flags, C.byref(raw), err, len(err))
        if rc != OK:
            raise Tp3Error(rc, "failed to load the configuration: " + err.value.decode())
        return cls(raw, flags)

    @classmethod
    def load(cls, file_name: str = "valeurs", features="") -> "Configuration":
        with open(file_name, "r") as f:
            return cls.parse(f.read(), features)

    @property
    def num_events(self) -> int:
        return int(self.raw.num_events)

    def with_num_events(self, n: int) -> "Configuration":
        raw = Config.from_buffer_copy(self.raw)
        raw.num_events = n
        return Configuration(raw, self.flags)

    def params(self, kernel: int = KERNEL_FAST) -> Params:
        """Couplings::new + EventGenerator::new + ResultsAccumulator::new -> kernel constants."""
        p = Params()
        rc = lib().tp3_params_from_config(C.byref(self.raw), self.flags, kernel, C.byref(p))
        if rc != OK:
            raise Tp3Error(rc, "tp3_params_from_config")
        return p


def merge(into: Acc, other: Acc, flags: int = 0) -> Acc:
    """ResultsAccumulator::merge (resacc.rs:133-139)."""
    lib().tp3_merge(C.byref(into), C.byref(other), flags)
    return into


def fold(accs: Sequence[Acc], flags: int = 0) -> Acc:
    """Left fold in batch order (sequential.rs:24-36, multi_threading.rs:107-126): tp3_fold_batches."""
    n = len(accs)
    arr = accs if isinstance(accs, C.Array) and accs._type_ is Acc else (Acc * n)(*accs)
    total = Acc()
    rc = lib().tp3_fold_batches(arr, n, flags, C.byref(total))
    if rc != OK:
        raise Tp3Error(rc, "tp3_fold_batches")
    return total


@dataclass
class FinalResults:
    """resfin.rs:26-62 plus the two text surfaces that are checked against the goldens."""
    raw: Final
    cfg: Configuration

    @property
    def selected_events(self) -> int:
        return int(self.raw.selected_events)

    @property
    def sigma(self) -> float:
        return float(self.raw.sigma)

    def res_data(self) -> str:
        n = lib().tp3_format_res_data(C.byref(self.cfg.raw), self.cfg.flags, C.byref(self.raw), None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().tp3_format_res_data(C.byref(self.cfg.raw), self.cfg.flags, C.byref(self.raw), buf, n + 1)
        return buf.value.decode()

    def stdout(self) -> str:
        n = lib().tp3_format_stdout(C.byref(self.cfg.raw), self.cfg.flags, C.byref(self.raw), None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().tp3_format_stdout(C.byref(self.cfg.raw), self.cfg.flags, C.byref(self.raw), buf, n + 1)
        return buf.value.decode()


def finalize(cfg: Configuration, merged: Acc) -> FinalResults:
    """ResultsAccumulator::finalize (resacc.rs:142-223)."""
    out = Final()
    rc = lib().tp3_finalize(C.byref(cfg.raw), cfg.flags, C.byref(merged), C.byref(out))
    if rc != OK:
        raise Tp3Error(rc, "tp3_finalize")
    return FinalResults(out, cfg)


def batch_layout(num_events: int):
    """(number of batches, length of the last one): multi_threading.rs:25,47."""
    nb = (num_events + EVENT_BATCH_SIZE - 1) // EVENT_BATCH_SIZE
    return nb, num_events - (nb - 1) * EVENT_BATCH_SIZE


def shard_range(n_batches: int, world_size: int, rank: int):
    """Contiguous batch range of `rank`: [n*rank/W, n*(rank+1)/W)."""
    lo = n_batches * rank // world_size
    hi = n_batches * (rank + 1) // world_size
    return lo, hi - lo


# ------------------------------------------------------------------------------ GPU context
class Simulator:
    """Owns a tp3_ctx: the B200 replacement for the `simulate_events` closure (main.rs:103-128),
    called for a RANGE of batches at a time."""

    def __init__(self, cfg: Configuration, kernel: int = KERNEL_FAST, devices: Optional[Sequence[int]] = None):
        self.cfg = cfg
        self.flags = cfg.flags
        self._params = cfg.params(kernel)
        devices = list(devices) if devices is not None else [0]
        ids = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = lib().tp3_create(C.byref(self._params), len(devices), ids, C.byref(h))
        if rc != OK:
            raise Tp3Error(rc, lib().tp3_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().tp3_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: the module globals may already be gone
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != OK:
            raise Tp3Error(rc, lib().tp3_last_error(self._h).decode())

    def set_stream(self, cuda_stream: int, slot: int = 0):
        self._check(lib().tp3_set_stream(self._h, slot, C.c_void_p(cuda_stream)))

    def simulate_batches(self, first_batch: int, n_batches: int, last_batch_len: int = EVENT_BATCH_SIZE):
        out = (Acc * n_batches)()
        self._check(lib().tp3_simulate_batches(self._h, first_batch, n_batches, last_batch_len, out))
        return out

    def simulate_batches_device(self, first_batch: int, n_batches: int, last_batch_len: int = EVENT_BATCH_SIZE):
        self._check(lib().tp3_simulate_batches_device(self._h, first_batch, n_batches, last_batch_len))

    def fetch(self, n_batches: int, out=None):
        out = out if out is not None else (Acc * n_batches)()
        self._check(lib().tp3_fetch(self._h, out, n_batches))
        return out

    def simulate_merged(self, first_batch: int, n_batches: int, last_batch_len: int = EVENT_BATCH_SIZE) -> Acc:
        out = Acc()
        self._check(lib().tp3_simulate_merged(self._h, first_batch, n_batches, last_batch_len, C.byref(out)))
        return out

    def simulate_batches_merged(self, first_batch: int, n_batches: int, last_batch_len: int = EVENT_BATCH_SIZE):
        """tp3_simulate_batches_merged: (one accumulator per batch on the host, their left fold in batch order done on the device)."""
        out = (Acc * n_batches)()
        merged = Acc()
        self._check(lib().tp3_simulate_batches_merged(self._h, first_batch, n_batches, last_batch_len, out, C.byref(merged)))
        return out, merged

    def simulate_merged_device(self, first_batch: int, n_batches: int, last_batch_len: int, device_ptr: int):
        """Asynchronous: the merged accumulator as 13 doubles in the caller's device buffer (one ncclReduce operand)."""
        self._check(lib().tp3_simulate_merged_device(self._h, first_batch, n_batches, last_batch_len, C.c_void_p(device_ptr)))

    def fe_tile_device(self, first_round: int, n_rounds: int, max_events: int, device_ptr: int) -> int:
        """faster-evgen stream tile (tp3_fe_tile_device): every event that starts in rounds [first_round, first_round +
        n_rounds), at most max_events; the merged accumulator goes to the device buffer, the event count is returned."""
        done = C.c_uint64()
        self._check(lib().tp3_fe_tile_device(self._h, first_round, n_rounds, max_events, C.c_void_p(device_ptr), C.byref(done)))
        return int(done.value)

    def synchronize(self):
        self._check(lib().tp3_synchronize(self._h))

    def set_option(self, name: str, value: int):
        """Test / A-B switches of include/tp3.h (tp3_set_option); nothing is read from the environment."""
        self._check(lib().tp3_set_option(self._h, name.encode(), int(value)))
        return self

    def get_stat(self, name: str) -> int:
        v = C.c_int64()
        self._check(lib().tp3_get_stat(self._h, name.encode(), C.byref(v)))
        return int(v.value)

    @property
    def launch_count(self) -> int:
        return int(lib().tp3_launch_count(self._h))

    # per-event observables: the hook the reference leaves empty (main.rs:117-122,133; include/tp3.h)
    def histograms_enable(self, num_bins: int):
        """Fill `num_bins`-bin histograms of x_k = 2 E_k / e_total and cos(theta_k) (k = 0, 1, 2) in every
        following simulate call; 0 switches them off."""
        self._check(lib().tp3_histograms_enable(self._h, num_bins))
        self._hist_bins = num_bins

    def histograms_reset(self):
        self._check(lib().tp3_histograms_reset(self._h))

    def histograms_fetch(self) -> "Histograms":
        n = HIST_OBSERVABLES * getattr(self, "_hist_bins", 0)
        counts = (C.c_uint64 * max(n, 1))()
        weights = (C.c_double * max(n, 1))()
        self._check(lib().tp3_histograms_fetch(self._h, counts, weights))
        nb = self._hist_bins
        return Histograms(nb, [list(counts[o * nb:(o + 1) * nb]) for o in range(HIST_OBSERVABLES)],
                          [list(weights[o * nb:(o + 1) * nb]) for o in range(HIST_OBSERVABLES)])

    def rng_dump(self, batch: int, n_words: int) -> List[int]:
        out = (C.c_uint64 * n_words)()
        self._check(lib().tp3_rng_dump(self._h, batch, n_words, out))
        return list(out)

    def events_dump(self, batch: int, n: int):
        mom = (C.c_double * (n * 12))()
        kept = (C.c_int32 * n)()
        m2 = (C.c_double * (n * 5))()
        self._check(lib().tp3_events_dump(self._h, batch, n, mom, kept, m2))
        return list(mom), list(kept), list(m2)

    def fastmath_probe(self, which: int, xs):
        n = len(xs)
        a = (C.c_double * n)(*xs)
        out = (C.c_double * n)()
        self._check(lib().tp3_fastmath_probe(self._h, which, n, a, out))
        return list(out)

    def peak_probe(self, which: int = 0) -> float:
        t = C.c_double()
        self._check(lib().tp3_peak_probe(self._h, which, C.byref(t)))
        return float(t.value)


HIST_OBSERVABLES = 6
HIST_NAMES = ("x_1", "x_2", "x_3", "cos_theta_1", "cos_theta_2", "cos_theta_3")
HIST_RANGES = ((0.0, 1.0),) * 3 + ((-1.0, 1.0),) * 3


class Histograms:
    """Per-event observables of the selected events: counts[o][b] events and weights[o][b] = sum of the event
    weights m . sigma_contribs (pb) in bin b of observable o (HIST_NAMES, uniform bins over HIST_RANGES).
    Additive over batches, devices and ranks (`+`)."""

    def __init__(self, num_bins: int, counts, weights):
        self.num_bins, self.counts, self.weights = num_bins, counts, weights

    def __add__(self, other: "Histograms") -> "Histograms":
        assert self.num_bins == other.num_bins
        return Histograms(self.num_bins, [[a + b for a, b in zip(x, y)] for x, y in zip(self.counts, other.counts)],
                          [[a + b for a, b in zip(x, y)] for x, y in zip(self.weights, other.weights)])

    def differential(self, o: int) -> List[float]:
        """d sigma / d observable per bin (what main.rs:133 calls normalising the histograms): weight / bin width."""
        lo, hi = HIST_RANGES[o]
        width = (hi - lo) / self.num_bins
        return [w / width for w in self.weights[o]]


def run_simulation(cfg: Configuration, kernel: int = KERNEL_FAST, devices: Optional[Sequence[int]] = None,
                   per_batch: bool = False) -> FinalResults:
    """scheduling::run_simulation (scheduling/mod.rs:31-59): all batches on the GPU(s), left fold in batch order
    (in the kernel: tp3_simulate_merged; `per_batch`: per-batch accumulators to the host and tp3_fold_batches there,
    the same bits), finalize."""
    nb, last = batch_layout(cfg.num_events)
    with Simulator(cfg, kernel, devices) as sim:
        total = fold(sim.simulate_batches(0, nb, last), cfg.flags) if per_batch else sim.simulate_merged(0, nb, last)
    return finalize(cfg, total)


def main_run(valeurs_path: str, out_dir: str = "", features="", kernel: int = KERNEL_FAST, n_dev: int = 1):
    """The whole program (main.rs:75-145) through tp3_run: returns (stdout text, elapsed seconds)."""
    buf = C.create_string_buffer(1 << 16)
    secs = C.c_double()
    rc = lib().tp3_run(valeurs_path.encode(), out_dir.encode(), feature_mask(features), kernel, n_dev, buf, len(buf), C.byref(secs))
    if rc != OK:
        raise Tp3Error(rc, buf.value.decode())
    return buf.value.decode(), float(secs.value)


def acc_from_f64x13(values) -> Acc:
    """The 13 doubles of tp3_simulate_merged_device (summed over ranks or not) back as an accumulator."""
    a = Acc()
    a.selected_events = int(round(values[0]))
    for k in range(5):
        a.spm2[k] = values[1 + k]
        a.vars[k] = values[6 + k]
    a.sigma, a.variance = values[11], values[12]
    return a


def acc_to_f64x13(a: Acc) -> List[float]:
    """An accumulator as the 13 doubles of tp3_simulate_merged_device: {count, spm2[5], vars[5], sigma, variance}."""
    return [float(a.selected_events)] + list(a.spm2) + list(a.vars) + [a.sigma, a.variance]


def run_simulation_reduced(cfg: Configuration, merged13, world_size: int, rank: int, dist=None):
    """scheduling::run_simulation over `world_size` processes with the ORDER-INSENSITIVE merge of the reference's
    faster-threading mode (FastAccumulator, multi_threading.rs:130-190): rank r simulates its contiguous batch range
    (shard_range; multi_threading.rs:25,46-70) and folds it in batch order; `merged13(first, n, last_len)` returns that
    fold as a torch tensor of 13 float64 (on the GPU: Simulator.simulate_merged_device into a CUDA tensor); ONE
    reduce(sum) to rank 0 is the run's only exchange; rank 0 finalizes (None elsewhere).  Reproducible for a fixed
    world size; selected_events is exact for any."""
    nb, last = batch_layout(cfg.num_events)
    lo, cnt = shard_range(nb, world_size, rank)
    my_last = last if lo + cnt == nb else EVENT_BATCH_SIZE
    t = merged13(lo, cnt, my_last)
    if world_size > 1:
        dist.reduce(t, dst=0)
        if rank != 0:
            return None
    return finalize(cfg, acc_from_f64x13(t.cpu().tolist()))


FE_ROUNDS_PER_EVENT = 0.325121  # rounds of 55 RANF numbers per faster-evgen event (measured over 2e9 events; sizes the tiles only)
FE_TILE_ALIGN = 512


def fe_tile_bounds(num_events: int, world_size: int, margin: float = 5e-4):
    """Round boundaries T_0 = 0 < T_1 < ... < T_W of the stream tiles of a `num_events`-event faster-evgen run: equal shares
    of the rounds the run is expected to need, less `margin` so that the last tile surely ends before event num_events
    (the last rank then adds the exact remainder)."""
    total = num_events * FE_ROUNDS_PER_EVENT * (1.0 - margin)
    return [int(total * r / world_size) // FE_TILE_ALIGN * FE_TILE_ALIGN for r in range(world_size + 1)]


def run_simulation_tiles(cfg: Configuration, tile13, world_size: int, rank: int, dist=None, device="cpu"):
    """scheduling::run_simulation for `faster-evgen` over `world_size` processes by sharding the STREAM instead of the
    batches (include/tp3.h, tp3_fe_tile_device).  `tile13(first_round, n_rounds, max_events) -> (tensor of 13 float64,
    events simulated)` is Simulator.fe_tile_device into a CUDA tensor on a GPU.  Exchanges: one all_reduce of the event
    counts (8 bytes), one reduce(sum) of 13 doubles.  Returns FinalResults on rank 0, None elsewhere."""
    import torch
    bounds = fe_tile_bounds(cfg.num_events, world_size)
    t, count = tile13(bounds[rank], bounds[rank + 1] - bounds[rank], 0)
    t = t.clone()
    total = torch.tensor([count], dtype=torch.int64, device=device)
    if world_size > 1:
        dist.all_reduce(total)
    remaining = cfg.num_events - int(total.item())
    if remaining < 0:
        raise Tp3Error(E_INVALID, f"faster-evgen tiles overshoot the run by {-remaining} events: FE_ROUNDS_PER_EVENT / margin need adjusting")
    if rank == world_size - 1 and remaining > 0:  # the exact remainder, from the end of the last tile
        t2, count2 = tile13(bounds[world_size], 0, remaining)
        if count2 != remaining:
            raise Tp3Error(E_INVALID, "faster-evgen remainder tile came back short")
        t = t + t2
    if world_size > 1:
        dist.reduce(t, dst=0)
        if rank != 0:
            return None
    return finalize(cfg, acc_from_f64x13(t.cpu().tolist()))


RUN_STAGES = ("read + parse valeurs", "context creation", "simulation", "finalize", "format + write outputs", "context destruction")


def main_run_stages(valeurs_path: str, out_dir: str = "", features="", kernel: int = KERNEL_FAST, n_dev: int = 1):
    """tp3_run_stages: (stdout text, reference-style elapsed seconds, {stage name: seconds})."""
    buf = C.create_string_buffer(1 << 16)
    secs = C.c_double()
    stages = (C.c_double * len(RUN_STAGES))()
    rc = lib().tp3_run_stages(valeurs_path.encode(), out_dir.encode(), feature_mask(features), kernel, n_dev, buf, len(buf),
                              C.byref(secs), stages)
    if rc != OK:
        raise Tp3Error(rc, buf.value.decode())
    return buf.value.decode(), float(secs.value), dict(zip(RUN_STAGES, stages))


# ------------------------------------------------------------------------------ multi-process
def gather_accumulators(local, world_size: int, rank: int, dist=None, device=None, dst: int = 0):
    """Per-batch accumulators of every rank, concatenated in rank (= batch) order on `dst`.
    `local` is this rank's ctypes array of Acc for its contiguous batch range (shard_range).  Works with
    any torch.distributed backend: byte tensors live on `device` ("cuda" for nccl, "cpu" for gloo).
    There is no data-path collective in the hot path; this is the one exchange at the end of a run
    (multi_threading.rs:107-126 gathers the same per-batch results from its worker threads)."""
    if world_size == 1:
        return list(local)
    import torch
    device = device or "cpu"
    n_local = len(local)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world_size)]
    dist.all_gather(counts, torch.tensor([n_local], dtype=torch.int64, device=device))
    counts = [int(c.item()) for c in counts]
    size = C.sizeof(Acc)
    width = max(counts) * size  # gather wants equal sizes: pad every rank's bytes to the longest range
    mine = torch.zeros(width, dtype=torch.uint8)
    if n_local:
        mine[: n_local * size] = torch.frombuffer(bytearray(bytes(local)), dtype=torch.uint8)
    mine = mine.to(device)
    parts = [torch.empty(width, dtype=torch.uint8, device=device) for _ in counts] if rank == dst else None
    dist.gather(mine, parts, dst=dst)
    if rank != dst:
        return None
    raw = b"".join(p.cpu().numpy().tobytes()[: c * size] for p, c in zip(parts, counts))
    return list((Acc * (len(raw) // size)).from_buffer_copy(raw))


def reduce_histograms(local: "Histograms", world_size: int, rank: int, dist=None, device=None, dst: int = 0):
    """Sum of every rank's per-event observable histograms on `dst` (None elsewhere): histograms are additive, so
    this is one small reduce at the end of a run, like the accumulator gather."""
    if world_size == 1:
        return local
    import torch
    device = device or "cpu"
    c = torch.tensor(local.counts, dtype=torch.int64, device=device)
    w = torch.tensor(local.weights, dtype=torch.float64, device=device)
    dist.reduce(c, dst)
    dist.reduce(w, dst)
    if rank != dst:
        return None
    return Histograms(local.num_bins, c.cpu().tolist(), w.cpu().tolist())


def run_simulation_distributed(cfg: Configuration, simulate_range, world_size: int, rank: int, dist=None, device=None):
    """scheduling::run_simulation over `world_size` processes (one GPU each): rank r simulates the
    contiguous batch range shard_range(r) with `simulate_range(first, n, last_len) -> Acc array`
    (Simulator.simulate_batches on a GPU), rank 0 folds all batches in batch order and finalizes.
    The result is bit-identical for every world_size."""
    nb, last = batch_layout(cfg.num_events)
    lo, cnt = shard_range(nb, world_size, rank)
    my_last = last if lo + cnt == nb else EVENT_BATCH_SIZE
    local = simulate_range(lo, cnt, my_last) if cnt else (Acc * 0)()
    accs = gather_accumulators(local, world_size, rank, dist, device)
    if accs is None:
        return None
    return finalize(cfg, fold(accs, cfg.flags))
