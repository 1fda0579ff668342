"""ctypes bindings for the CPU oracles - TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``traceit_b200``) never does.

Two libraries:

* ``oracle/_build/liboracle.so``  - restated C oracle (``oracle/restated.c``), always
  buildable with gcc (``make -C oracle``).
* ``oracle/_ref/libtraceit_ref.so`` - the UNMODIFIED reference translation unit
  (``/root/reference/src/main.cpp``) compiled in place behind a stub ncurses header
  (``make -C oracle ref``); present only when it was built where the reference exists.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libtraceit_ref.so")
REF_O0_SO = os.path.join(HERE, "_ref", "libtraceit_ref_O0.so")

# Shared record layout: hot fields of the reference's object + collider
# (reference src/main.cpp:79-105).  92 bytes.
BODY_DTYPE = np.dtype(
    [
        ("pos", "<f4", 3),
        ("vel", "<f4", 3),
        ("force", "<f4", 3),
        ("mass", "<f4"),
        ("Cd", "<f4"),
        ("area", "<f4"),
        ("stationary", "<i4"),
        ("center", "<f4", 3),
        ("size", "<f4", 3),
        ("radius", "<f4"),
        ("isSphere", "<i4"),
        ("isStatic", "<i4"),
        ("rest", "<f4"),
    ],
    align=False,
)
assert BODY_DTYPE.itemsize == 92

PAIR_GRAVITY = 1
COLLISIONS = 2
GRAVITY_F64 = 4


class Params(C.Structure):
    _fields_ = [
        ("g", C.c_float * 3),
        ("G", C.c_float),
        ("dt", C.c_float),
        ("atm_density", C.c_float),
        ("flags", C.c_int32),
    ]


def build(ref: bool = True) -> None:
    """Compile the oracle libraries (restated always; verbatim when the reference exists)."""
    subprocess.run(["make", "-C", HERE, "-s"], check=True)
    if ref and os.path.exists("/root/reference/src/main.cpp"):
        subprocess.run(["make", "-C", HERE, "-s", "ref"], check=True)


_lib = None
_ref = {}


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = C.CDLL(ORACLE_SO)
        L.orc_default_params.argtypes = [C.POINTER(Params)]
        L.orc_step.argtypes = [C.c_void_p, C.c_int, C.POINTER(Params), C.c_int]
        L.orc_integrate.argtypes = [C.c_void_p, C.c_int, C.POINTER(Params), C.c_void_p]
        L.orc_resolve.argtypes = [C.c_void_p, C.c_int]
        L.orc_resolve_list.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
        L.orc_pair_gravity_f32.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        L.orc_pair_gravity_f64.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.orc_pair_gravity_f64_subset.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_contacts.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
        L.orc_contacts.restype = C.c_longlong
        L.orc_grid_cells.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p]
        L.orc_grid_keys.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        L.orc_grid_in_domain.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p]
        L.orc_grid_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_grid_pairs.argtypes = [
            C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int,
            C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong),
            C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong),
        ]
        L.orc_grid_pairs.restype = C.c_int
        L.orc_state_hash.argtypes = [C.c_void_p, C.c_int]
        L.orc_state_hash.restype = C.c_uint64
        L.orc_sizeof_body.restype = C.c_int
        assert L.orc_sizeof_body() == BODY_DTYPE.itemsize
        _lib = L
    return _lib


def default_params(**over) -> Params:
    p = Params()
    lib().orc_default_params(C.byref(p))
    for k, v in over.items():
        if k == "g":
            p.g[0], p.g[1], p.g[2] = [float(x) for x in v]
        else:
            setattr(p, k, v)
    return p


def new_bodies(n: int) -> np.ndarray:
    """Zeroed body records with the UI defaults that matter (Cd 0.42, rest 0.5)."""
    b = np.zeros(n, dtype=BODY_DTYPE)
    b["Cd"] = np.float32(0.42)
    b["rest"] = np.float32(0.5)
    b["mass"] = np.float32(1.0)
    return b


# ---------------------------------------------------------------- restated oracle
def step(bodies: np.ndarray, params: Params, nsteps: int = 1) -> np.ndarray:
    assert bodies.dtype == BODY_DTYPE and bodies.flags.c_contiguous
    lib().orc_step(_ptr(bodies), len(bodies), C.byref(params), nsteps)
    return bodies


def integrate(bodies: np.ndarray, params: Params, fgrav: np.ndarray | None = None) -> np.ndarray:
    if fgrav is not None:
        fgrav = np.ascontiguousarray(fgrav, dtype=np.float32)
    lib().orc_integrate(_ptr(bodies), len(bodies), C.byref(params), _ptr(fgrav) if fgrav is not None else None)
    return bodies


def resolve(bodies: np.ndarray) -> np.ndarray:
    lib().orc_resolve(_ptr(bodies), len(bodies))
    return bodies


def resolve_list(bodies: np.ndarray, pairs: np.ndarray) -> np.ndarray:
    pairs = np.ascontiguousarray(pairs, dtype=np.int32)
    lib().orc_resolve_list(_ptr(bodies), _ptr(pairs), len(pairs))
    return bodies


def pair_gravity_f32(bodies: np.ndarray, G: float) -> np.ndarray:
    out = np.empty((len(bodies), 3), dtype=np.float32)
    lib().orc_pair_gravity_f32(_ptr(bodies), len(bodies), C.c_float(G), _ptr(out))
    return out


def pair_gravity_f64(bodies: np.ndarray, G: float) -> np.ndarray:
    out = np.empty((len(bodies), 3), dtype=np.float64)
    lib().orc_pair_gravity_f64(_ptr(bodies), len(bodies), C.c_double(G), _ptr(out))
    return out


def pair_gravity_f64_subset(bodies: np.ndarray, G: float, idx) -> np.ndarray:
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.empty((len(idx), 3), dtype=np.float64)
    lib().orc_pair_gravity_f64_subset(_ptr(bodies), len(bodies), C.c_double(G), _ptr(idx), len(idx), _ptr(out))
    return out


def contacts(bodies: np.ndarray) -> np.ndarray:
    """Brute-force contact set (lexicographic (i,j) rows)."""
    n = lib().orc_contacts(_ptr(bodies), len(bodies), None, 0)
    out = np.empty((n, 2), dtype=np.int32)
    if n:
        lib().orc_contacts(_ptr(bodies), len(bodies), _ptr(out), n)
    return out


def grid_keys(bodies: np.ndarray, origin, inv_cell: float, bits: int) -> np.ndarray:
    o = np.asarray(origin, dtype=np.float32)
    keys = np.empty(len(bodies), dtype=np.uint32)
    lib().orc_grid_keys(_ptr(bodies), len(bodies), _ptr(o), C.c_float(inv_cell), bits, _ptr(keys))
    return keys


def grid_cells(bodies: np.ndarray, origin, inv_cell: float) -> np.ndarray:
    o = np.asarray(origin, dtype=np.float32)
    cells = np.empty((len(bodies), 3), dtype=np.int32)
    lib().orc_grid_cells(_ptr(bodies), len(bodies), _ptr(o), C.c_float(inv_cell), _ptr(cells))
    return cells


def grid_in_domain(bodies: np.ndarray, origin, inv_cell: float) -> np.ndarray:
    o = np.asarray(origin, dtype=np.float32)
    ok = np.empty(len(bodies), dtype=np.uint8)
    lib().orc_grid_in_domain(_ptr(bodies), len(bodies), _ptr(o), C.c_float(inv_cell), _ptr(ok))
    return ok


def grid_sort(keys: np.ndarray, ids: np.ndarray, bits: int):
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    order = np.empty(len(ids), dtype=np.int32)
    skeys = np.empty(len(ids), dtype=np.uint32)
    lib().orc_grid_sort(_ptr(keys), _ptr(ids), len(ids), bits, _ptr(order), _ptr(skeys))
    return order, skeys


def grid_pairs(bodies, in_grid, origin, inv_cell, bits, want_candidates=True, want_contacts=True):
    """(candidates, contacts) through the CPU grid restatement; each an (m,2) int32 array."""
    o = np.asarray(origin, dtype=np.float32)
    in_grid = np.ascontiguousarray(in_grid, dtype=np.uint8)
    nc = C.c_longlong(0)
    nk = C.c_longlong(0)
    rc = lib().orc_grid_pairs(_ptr(bodies), len(bodies), _ptr(in_grid), _ptr(o), C.c_float(inv_cell), bits,
                              None, 0, C.byref(nc), None, 0, C.byref(nk))
    assert rc == 0
    cand = np.empty((nc.value if want_candidates else 0, 2), dtype=np.int32)
    cont = np.empty((nk.value if want_contacts else 0, 2), dtype=np.int32)
    rc = lib().orc_grid_pairs(_ptr(bodies), len(bodies), _ptr(in_grid), _ptr(o), C.c_float(inv_cell), bits,
                              _ptr(cand) if want_candidates and nc.value else None, len(cand), C.byref(nc),
                              _ptr(cont) if want_contacts and nk.value else None, len(cont), C.byref(nk))
    assert rc == 0
    return cand, cont, nc.value, nk.value


def state_hash(bodies: np.ndarray) -> int:
    return int(lib().orc_state_hash(_ptr(bodies), len(bodies)))


# ---------------------------------------------------------------- verbatim reference
def ref_available(o0: bool = False) -> bool:
    return os.path.exists(REF_O0_SO if o0 else REF_SO)


def reflib(o0: bool = False):
    key = "O0" if o0 else "O2"
    if key not in _ref:
        path = REF_O0_SO if o0 else REF_SO
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        # RTLD_LOCAL (ctypes default) keeps the reference TU's many globals private.
        L = C.CDLL(path)
        L.ref_world_create.argtypes = [C.c_float, C.c_char_p]
        L.ref_world_create.restype = C.c_void_p
        L.ref_world_destroy.argtypes = [C.c_void_p]
        L.ref_world_size.argtypes = [C.c_void_p]
        L.ref_world_set_atm.argtypes = [C.c_void_p, C.c_float]
        L.ref_world_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ref_world_step.argtypes = [C.c_void_p, C.c_int]
        L.ref_world_resolve.argtypes = [C.c_void_p]
        L.ref_world_read.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_world_set_force.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
        L.ref_world_contacts.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
        L.ref_world_contacts.restype = C.c_longlong
        L.ref_to_spherical.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.ref_area_from_size.argtypes = [C.c_float]
        L.ref_area_from_size.restype = C.c_float
        L.ref_sizeof_object.restype = C.c_int
        _ref[key] = L
    return _ref[key]


class RefWorld:
    """The reference's own `world` + `updatePhysics`, driven headless."""

    def __init__(self, atm_density: float = 1.2, o0: bool = False):
        self.L = reflib(o0)
        self._tmp = tempfile.TemporaryDirectory(prefix="traceit_ref_")
        self.h = self.L.ref_world_create(C.c_float(atm_density), self._tmp.name.encode())

    def add(self, bodies: np.ndarray) -> int:
        assert bodies.dtype == BODY_DTYPE
        bodies = np.ascontiguousarray(bodies)
        return self.L.ref_world_add(self.h, _ptr(bodies), len(bodies))

    def step(self, nsteps: int = 1) -> None:
        self.L.ref_world_step(self.h, nsteps)

    def resolve(self) -> None:
        self.L.ref_world_resolve(self.h)

    def set_force(self, idx: int, f) -> None:
        self.L.ref_world_set_force(self.h, idx, float(f[0]), float(f[1]), float(f[2]))

    def read(self) -> np.ndarray:
        n = self.L.ref_world_size(self.h)
        out = np.zeros(n, dtype=BODY_DTYPE)
        self.L.ref_world_read(self.h, _ptr(out), 0, n)
        return out

    def contacts(self) -> np.ndarray:
        n = self.L.ref_world_contacts(self.h, None, 0)
        out = np.empty((n, 2), dtype=np.int32)
        if n:
            self.L.ref_world_contacts(self.h, _ptr(out), n)
        return out

    def close(self) -> None:
        if self.h:
            self.L.ref_world_destroy(self.h)
            self.h = None
            self._tmp.cleanup()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_to_spherical(elev: float, azim: float, magn: float) -> np.ndarray:
    out = np.zeros(3, dtype=np.float32)
    reflib().ref_to_spherical(C.c_float(elev), C.c_float(azim), C.c_float(magn), _ptr(out))
    return out


def ref_area_from_size(size: float) -> np.float32:
    return np.float32(reflib().ref_area_from_size(C.c_float(size)))
