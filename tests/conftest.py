import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-exact float32 equality; any NaN equals any NaN (payload/sign not compared)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    if a.shape != b.shape:
        return False
    same = a.view(np.uint32) == b.view(np.uint32)
    return bool(np.all(same | (np.isnan(a) & np.isnan(b))))


def mismatch_report(a, b, name=""):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    bad = ~((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)))
    n = int(bad.sum())
    if n == 0:
        return f"{name}: identical"
    ys, xs = np.nonzero(bad)
    first = [(int(x), int(y), float(a[y, x]), float(b[y, x])) for y, x in list(zip(ys, xs))[:6]]
    with np.errstate(all="ignore"):
        rel = np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-30))
    return f"{name}: {n} mismatching cells of {a.size}; max rel {rel:.3e}; first (x,y,got,want): {first}"


def assert_bits(a, b, name=""):
    assert bits_equal(a, b), mismatch_report(a, b, name)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def rand_field(rng, w, h, scale=1.0):
    return (rng.standard_normal((h, w)) * scale).astype(np.float32)


@pytest.fixture(scope="session")
def fb():
    """The product package over libfruid_b200.so (built in-tree; raises if the library is missing)."""
    import __graft_entry__
    __graft_entry__.build()
    import fruid_b200
    return fruid_b200
