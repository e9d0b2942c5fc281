"""Test fixtures: the product package (3photons-rust_b200, loaded by path because its directory
name is not an identifier), the CPU oracle (tests' checker) and the reference's golden files."""
import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


def load_package():
    name = "tp3b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "3photons-rust_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _ensure_built():
    """Build the product library and the oracle if a fresh checkout has neither (CPU only: nvcc
    cross-compiles). On the GPU box the prebuilt files travel with the snapshot."""
    lib = os.path.join(ROOT, "3photons-rust_b200", "_build", "libtp3.so")
    orc = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "3photons-rust_b200")])
    if not os.path.exists(orc):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])


@pytest.fixture(scope="session")
def tp3():
    _ensure_built()
    return load_package()


@pytest.fixture(scope="session")
def oracle():
    _ensure_built()
    import oracle_lib
    return oracle_lib


@pytest.fixture(scope="session")
def valeurs_text():
    with open(os.path.join(GOLDEN, "valeurs")) as f:
        return f.read()


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return f.read()
