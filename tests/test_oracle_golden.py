"""Pins the CPU oracle against every golden file of the reference (reference/res.data-* and
reference/stdout.log-*, copied to tests/golden/), with the tolerances of the reference's own CI
(.github/workflows/ci.yml:179-203): exact, except f32 (res.data abs 1.1e-8, stdout rel 1.9e-5)
and standard-random,multi-threading,faster-threading (abs 1.1e-16, racy merge order)."""
import pytest

from conftest import golden
from numdiff import compare

# (features, golden suffix): the symlinks of reference/ are spelled out
CASES = [
    ("", ""),
    ("f32", "f32"),
    ("faster-evgen", "faster-evgen"),
    ("multi-threading", ""),
    ("multi-threading,faster-threading", "multi-threading,faster-threading"),
    ("no-photon-sorting", ""),
    ("standard-random", "standard-random"),
    ("standard-random,f32", "standard-random,f32"),
    ("standard-random,multi-threading", "standard-random"),
    ("standard-random,multi-threading,faster-threading", "standard-random,multi-threading,faster-threading"),
]


def tolerances(features):
    if "f32" in features:
        return dict(res=dict(abs_=1.1e-8), out=dict(rel=1.9e-5))
    if features == "standard-random,multi-threading,faster-threading":
        return dict(res=dict(abs_=1.1e-16), out=dict())
    return dict(res=dict(), out=dict())


@pytest.mark.parametrize("features,suffix", CASES, ids=[c[0] or "default" for c in CASES])
def test_oracle_reproduces_golden(oracle, valeurs_text, features, suffix):
    run = oracle.run(valeurs_text, features, threads=8, want_batches=False)
    tol = tolerances(features)
    assert compare(run.res_data, golden("res.data-features_" + suffix), **tol["res"]) == []
    assert compare(run.stdout, golden("stdout.log-features_" + suffix), **tol["out"]) == []
