"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/traceit_b200.h declares, has the documented struct layouts and reference-constant
defaults, and FAILS LOUDLY (no CPU fallback) when asked to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import traceit_b200 as tr
from conftest import ROOT, assert_bits_equal, load_golden

HEADER = os.path.join(ROOT, "include", "traceit_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tr_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_loads():
    assert os.path.exists(tr.lib_path()), "run __graft_entry__.build() first"
    assert tr.lib().tr_abi_version() == 1


def test_every_declared_symbol_is_exported():
    declared = _declared_symbols()
    assert len(declared) >= 30
    L = C.CDLL(tr.lib_path())
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(tr.ABI_SYMBOLS) == declared, "python binding list out of sync with the header"


def test_struct_layouts_match_header():
    assert tr.BODY_DTYPE.itemsize == 92
    assert C.sizeof(tr.Params) == 72
    assert tr.Params.max_contacts.offset == 48
    assert C.sizeof(tr.Stats) == 128   # (+ sparse_steps, sparse_bailouts)
    # same record as the oracle's (tests pass arrays between the two)
    from oracle import pyoracle as po
    assert po.BODY_DTYPE == tr.BODY_DTYPE


def test_default_params_are_the_reference_constants():
    p = tr.default_params()
    assert_bits_equal(np.array(list(p.g), np.float32), np.array([0, np.float32(-9.8), 0], np.float32))  # main.cpp:749
    assert np.float32(p.dt) == np.float32(0.01666)          # main.cpp:750
    assert np.float32(p.G) == np.float32(6.6743e-11)        # main.cpp:685
    assert np.float32(p.atm_density) == np.float32(1.2)     # main.cpp:1141
    assert p.flags == tr.COLLISIONS                         # pair gravity is dead code in the reference


def test_creation_helpers_match_reference_goldens():
    g = load_golden("creation_helpers.npz")
    for a, want in zip(g["angles"], g["spherical"]):
        assert_bits_equal(tr.to_spherical(float(a[0]), float(a[1]), float(a[2])), want, f"toSpherical{tuple(a)}")
    for s, want in zip(g["sizes"], g["areas"]):
        assert_bits_equal(tr.area_from_size(float(s)), want, f"area({s})")


def test_no_cpu_fallback():
    n = C.c_int(-1)
    rc = tr.lib().tr_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present; the loud-failure path is for CPU-only hosts")
    with pytest.raises(tr.TraceItError) as e:
        tr.World()
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    with pytest.raises(tr.TraceItError):
        tr.measure_fma_peak(0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under traceit_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "traceit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"\bpyoracle\b|liboracle|oracle/|import oracle|from oracle", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_header_is_plain_c99_and_links(tmp_path):
    """include/traceit_b200.h compiles as C99 and every entry point links against the .so."""
    import subprocess
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    src = os.path.join(ROOT, "tests", "c_abi_check.c")
    out = tmp_path / "c_abi_check.so"
    cmd = [cc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), src,
           "-L", os.path.dirname(tr.lib_path()), "-ltraceit_b200", "-Wl,--no-undefined", "-Wl,--allow-shlib-undefined",
           "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    used = set(re.findall(r"\b(tr_[a-z0-9_]+)\s*\(", open(src).read()))
    assert set(_declared_symbols()) <= used, sorted(set(_declared_symbols()) - used)


def test_minimal_c_example_builds(tmp_path):
    import subprocess
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    out = tmp_path / "minimal"
    libdir = os.path.dirname(tr.lib_path())
    cmd = [cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "minimal.c"),
           "-L", libdir, "-ltraceit_b200", f"-Wl,-rpath,{libdir}", "-Wl,-rpath-link,/usr/local/cuda/lib64", "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    n = C.c_int(0)
    if tr.lib().tr_device_count(C.byref(n)) != 0 or n.value == 0:
        run = subprocess.run([str(out)], capture_output=True, text=True)
        assert run.returncode == 1 and "no CPU fallback" in run.stderr      # loud failure without a GPU
