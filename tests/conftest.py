import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # build the native pieces on demand (nvcc cross-compiles without a GPU); the GPU box normally
    # receives them prebuilt with the snapshot
    need = [os.path.join(ROOT, "traceit_b200", "libtraceit_b200.so"), os.path.join(ROOT, "oracle", "_build", "liboracle.so"),
            os.path.join(ROOT, "traceit_b200", "host", "traceit_headless")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


def bits(a):
    """Bit pattern view of a float32 array (for exact comparisons incl. signed zeros / NaNs)."""
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits_equal(a, b, what=""):
    """Exact bit equality (signed zeros and infinities included).  NaNs compare equal to
    NaNs: IEEE 754 leaves the payload of a generated NaN unspecified and the platforms differ
    (x86 SSE produces 0xFFC00000, NVIDIA GPUs 0x7FFFFFFF)."""
    fa, fb = np.ascontiguousarray(a, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)
    a, b = bits(fa).copy(), bits(fb).copy()
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    a[np.isnan(fa)] = 0x7FC00000
    b[np.isnan(fb)] = 0x7FC00000
    bad = np.nonzero(a != b)
    assert len(bad[0]) == 0, f"{what}: {len(bad[0])} of {a.size} words differ, first at {tuple(x[0] for x in bad)}"


def mixed_cloud():
    """Same construction as tests/golden/make_golden.py::mixed_cloud."""
    from traceit_b200 import scenes
    m = scenes.random_cloud(1024, 77, spheres=True, density_l=0.5)
    m["isSphere"][::3] = 0
    m["size"][::3] = (0.6, 0.4, 0.5)
    m["isStatic"][::7] = 1
    m["stationary"][::11] = 1
    m["mass"][::13] = 0
    m["rest"][::5] = 0.9
    return m


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle as po
    po.lib()
    return po


@pytest.fixture(scope="session")
def tr():
    """The product binding; gpu tests fail (not skip) when the extension is missing."""
    import traceit_b200
    traceit_b200.lib()
    return traceit_b200


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))
