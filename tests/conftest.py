import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _bits(a):
    """float32 bit patterns with every NaN mapped to one canonical pattern: which NaN a 0/0 produces (x86 gives the
    negative quiet NaN, the GPU 0x7fffffff) is not observable through the engine's API; NaN-ness is."""
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).copy()
    u[np.isnan(a)] = 0x7fc00000
    return u


def biteq(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(_bits(a), _bits(b))


def assert_biteq(a, b, what=""):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    bad = np.argwhere(_bits(a) != _bits(b))
    assert len(bad) == 0, "%s: %d of %d floats differ, first at %s: got %r want %r" % (
        what, len(bad), a.size, tuple(bad[0]), a[tuple(bad[0])], b[tuple(bad[0])])


def gold(name):
    return np.load(os.path.join(GOLD, name))


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build what is missing or stale (CUDA library via nvcc, C oracle via gcc) before any test touches it."""
    import __graft_entry__ as entry
    try:
        entry.build()
    except Exception as e:                      # a box without nvcc still runs with the prebuilt library that travelled with the repo
        if not os.path.exists(entry.LIB):
            raise
        sys.stderr.write("build() skipped: %s\n" % e)


@pytest.fixture(scope="session")
def oracle():
    from oracle.bind import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own engine, compiled in the dev container only (absent on the GPU box unless prebuilt)."""
    from oracle.bind import Ref, ref_available
    if not ref_available() and not os.path.isdir("/root/reference/engine"):
        pytest.skip("oracle/_ref/libnetphys_ref.so not built and /root/reference absent")
    return Ref()


@pytest.fixture(scope="session")
def npb():
    import netphys_b200
    return netphys_b200


@pytest.fixture()
def world(npb):
    if npb.device_count() < 1:
        pytest.fail("gpu test selected but no CUDA device is visible (netphys_b200 has no CPU path)")
    w = npb.World(0)
    yield w
    w.close()
