import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _load(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def kat():
    d = _load("kat.npz")
    d["meta"] = json.loads(str(d["meta"]))
    return d


@pytest.fixture(scope="session")
def stats_golden():
    d = _load("stats.npz")
    d["meta"] = json.loads(str(d["meta"]))
    return d


@pytest.fixture(scope="session")
def full_golden():
    """BASELINE configs C1 / C2 at their stated size (1,000 paths x 1,000 trades) captured from the live reference;
    the sequences are stored as a value table + uint8 codes and expanded here."""
    d = _load("full.npz")
    d["meta"] = json.loads(str(d["meta"]))
    for name in d["meta"]:
        d[name + "/seqs"] = np.ascontiguousarray(d[name + "/values"][d[name + "/codes"]])
    return d


@pytest.fixture(scope="session")
def rng_golden():
    return _load("rng.npz")


def bits(a):
    """float64 array -> uint64 bit patterns (bit-exact comparisons, NaN/zero-sign safe)."""
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bitexact(a, b, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    bad = np.flatnonzero(bits(a).ravel() != bits(b).ravel())
    assert bad.size == 0, "%s: %d of %d values differ, first at %d: %r vs %r" % (
        what, bad.size, a.size, bad[0], a.ravel()[bad[0]], b.ravel()[bad[0]])
