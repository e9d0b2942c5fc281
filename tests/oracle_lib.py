"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_build", "liboracle.so")

F32, FASTER_EVGEN, FASTER_THREADING, MULTI_THREADING, NO_PHOTON_SORTING, STANDARD_RANDOM = 1, 2, 4, 8, 16, 32
BITS = {"f32": F32, "faster-evgen": FASTER_EVGEN, "faster-threading": FASTER_THREADING,
        "multi-threading": MULTI_THREADING, "no-photon-sorting": NO_PHOTON_SORTING, "standard-random": STANDARD_RANDOM}


def mask(features):
    if isinstance(features, int):
        return features
    m = 0
    for f in (features.split(",") if isinstance(features, str) else features):
        if f:
            m |= BITS[f]
    return m


class Acc(C.Structure):
    _fields_ = [("selected_events", C.c_uint64), ("spm2", C.c_double * 5), ("vars", C.c_double * 5),
                ("sigma", C.c_double), ("variance", C.c_double)]


_lib = None


def use_library(path):
    """Load another build of the same oracle (bench.py's -O3 -march=native rebuild on the GPU box)."""
    global _lib, LIB_PATH
    LIB_PATH = path
    _lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.oracle_run.restype = C.c_int
        L.oracle_run.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_uint64, C.POINTER(Acc), C.POINTER(C.c_uint64),
                                 C.POINTER(Acc), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_double)]
        L.oracle_rng_words.restype = C.c_int
        L.oracle_rng_words.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]
        L.oracle_events.restype = C.c_int
        L.oracle_events.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_double)]
        L.oracle_finalize_text.restype = C.c_int
        L.oracle_finalize_text.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(Acc), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.oracle_fe_tile.restype = C.c_int
        L.oracle_fe_tile.argtypes = [C.c_char_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(Acc), C.POINTER(C.c_uint64)]
        L.oracle_constants.restype = C.c_int
        L.oracle_constants.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_double)]
        _lib = L
    return _lib


class Run:
    def __init__(self, per_batch, merged, res_data, stdout, seconds):
        self.per_batch, self.merged, self.res_data, self.stdout, self.seconds = per_batch, merged, res_data, stdout, seconds


def run(valeurs_text, features="", threads=1, num_events=0, want_batches=True, want_text=True):
    """Whole run through the oracle. num_events=0 keeps the file's value."""
    m = mask(features)
    cap = 0
    per_batch = None
    if want_batches:
        n = num_events or int(valeurs_text.split()[0])
        cap = n // 10000 + 2
        per_batch = (Acc * cap)()
    nb = C.c_uint64(cap)
    merged = Acc()
    rd = C.create_string_buffer(1 << 14) if want_text else None
    so = C.create_string_buffer(1 << 14) if want_text else None
    secs = C.c_double()
    rc = lib().oracle_run(valeurs_text.encode(), m, threads, num_events, per_batch, C.byref(nb), C.byref(merged),
                          rd, len(rd) if rd else 0, so, len(so) if so else 0, C.byref(secs))
    if rc != 0:
        raise RuntimeError("oracle_run failed: " + (so.value.decode() if so else ""))
    batches = [per_batch[i] for i in range(nb.value)] if want_batches else None
    return Run(batches, merged, rd.value.decode() if rd else None, so.value.decode() if so else None, secs.value)


def rng_words(features, batch, n_words):
    out = (C.c_uint64 * n_words)()
    lib().oracle_rng_words(mask(features), batch, n_words, out)
    return list(out)


def events(valeurs_text, features, n):
    mom = (C.c_double * (n * 12))()
    kept = (C.c_int32 * n)()
    m2 = (C.c_double * (n * 5))()
    rc = lib().oracle_events(valeurs_text.encode(), mask(features), n, mom, kept, m2)
    assert rc == 0
    return list(mom), list(kept), list(m2)


def fe_tile(valeurs_text, features, first_round, n_rounds, max_events):
    """(merged accumulator, number of events) of the faster-evgen events that start in RANF rounds
    [first_round, first_round + n_rounds), at most max_events of them (0 = no limit): the checker of tp3_fe_tile_device."""
    acc = Acc()
    done = C.c_uint64()
    rc = lib().oracle_fe_tile(valeurs_text.encode(), mask(features), first_round, n_rounds, max_events, C.byref(acc), C.byref(done))
    assert rc == 0, rc
    return acc, int(done.value)


def constants(valeurs_text, features=""):
    out = (C.c_double * 10)()
    rc = lib().oracle_constants(valeurs_text.encode(), mask(features), out)
    assert rc == 0
    return list(out)


def finalize_text(valeurs_text, features, acc):
    """(res.data text, stdout text) of the oracle's finalize + writers applied to the given sums (an Acc)."""
    rd = C.create_string_buffer(1 << 14)
    so = C.create_string_buffer(1 << 14)
    mine = Acc.from_buffer_copy(bytes(acc))  # same 104-byte layout as the product's tp3_acc
    rc = lib().oracle_finalize_text(valeurs_text.encode(), mask(features), C.byref(mine), rd, len(rd), so, len(so))
    assert rc == 0
    return rd.value.decode(), so.value.decode()
