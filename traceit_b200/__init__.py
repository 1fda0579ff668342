"""traceit_b200 - B200-native TraceIt physics step (host-side Python binding).

The product is ``libtraceit_b200.so`` (hand-written CUDA for sm_100a behind the C ABI in
``include/traceit_b200.h``).  This module is a thin ctypes mirror of that ABI used by the
tests and ``bench.py``; the reference-shaped C++ host shim lives in
``traceit_b200/host/traceit_host.hpp``.

There is no CPU fallback: importing works anywhere (so the ABI can be inspected), but
every compute call raises :class:`TraceItError` when the extension or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

__all__ = [
    "BODY_DTYPE", "Params", "Stats", "TraceItError", "World", "lib", "lib_path", "build_extension",
    "default_params", "to_spherical", "area_from_size", "new_bodies", "PAIR_GRAVITY", "COLLISIONS",
    "comm_unique_id", "measure_fma_peak", "ABI_SYMBOLS", "FRAME_DTYPE",
]

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

PAIR_GRAVITY = 1
COLLISIONS = 2
DEBUG_NO_TINY = 1
DEBUG_BIG_LIST = 2
DEBUG_NO_SPARSE = 4

ERRORS = {
    0: "TR_OK", 1: "TR_ERR_INVALID_ARG", 2: "TR_ERR_NO_DEVICE", 3: "TR_ERR_CUDA", 4: "TR_ERR_NCCL",
    5: "TR_ERR_CAPACITY", 6: "TR_ERR_DOMAIN", 7: "TR_ERR_STATE", 8: "TR_ERR_NOMEM",
}

# tr_body_desc: hot fields of the reference's object + collider (reference src/main.cpp:79-105)
BODY_DTYPE = np.dtype(
    [
        ("pos", "<f4", 3), ("vel", "<f4", 3), ("force", "<f4", 3),
        ("mass", "<f4"), ("Cd", "<f4"), ("area", "<f4"), ("stationary", "<i4"),
        ("center", "<f4", 3), ("size", "<f4", 3), ("radius", "<f4"),
        ("isSphere", "<i4"), ("isStatic", "<i4"), ("rest", "<f4"),
    ],
    align=False,
)
assert BODY_DTYPE.itemsize == 92


class Params(C.Structure):
    """tr_world_params (defaults = the reference's constants, see tr_default_params)."""

    _fields_ = [
        ("g", C.c_float * 3), ("G", C.c_float), ("dt", C.c_float), ("atm_density", C.c_float),
        ("flags", C.c_uint32), ("grid_origin", C.c_float * 3), ("grid_cell", C.c_float),
        ("grid_bits", C.c_int32), ("max_contacts", C.c_uint64), ("grid_max_radius", C.c_float),
        ("response_mode", C.c_uint32), ("debug_flags", C.c_uint32), ("reserved", C.c_uint32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("steps", C.c_uint64), ("kernel_launches", C.c_uint64), ("last_contacts", C.c_uint64),
        ("last_candidates", C.c_uint64), ("last_response_rounds", C.c_uint32), ("side_list_size", C.c_uint32),
        ("gravity_ms", C.c_double), ("gravity_launches", C.c_uint64), ("broadphase_ms", C.c_double),
        ("broadphase_launches", C.c_uint64), ("narrowphase_ms", C.c_double), ("response_ms", C.c_double),
        ("integrate_ms", C.c_double), ("allgather_ms", C.c_double), ("interactions", C.c_uint64),
        ("sparse_steps", C.c_uint64), ("sparse_bailouts", C.c_uint64),
    ]


class TraceItError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


# every symbol include/traceit_b200.h declares (tests check the .so exports all of them)
ABI_SYMBOLS = [
    "tr_abi_version", "tr_last_error", "tr_device_count", "tr_default_params", "tr_world_create",
    "tr_comm_unique_id", "tr_world_create_sharded", "tr_world_destroy", "tr_world_set_params",
    "tr_world_get_params", "tr_world_set_stream", "tr_body_add", "tr_bodies_add", "tr_world_size",
    "tr_world_owned_range", "tr_to_spherical", "tr_area_from_size", "tr_step", "tr_sync", "tr_step_host",
    "tr_read_state", "tr_read_bodies", "tr_write_bodies", "tr_set_force", "tr_read_contacts", "tr_debug_grid",
    "tr_debug_candidates", "tr_debug_gravity", "tr_get_stats", "tr_reset_stats", "tr_enable_timing",
    "tr_measure_fma_peak", "tr_world_enable_trajectory", "tr_read_trajectory", "tr_step_frame",
]

# tr_frame_body: what one step changes (pos, vel, collider centre), 36 bytes
FRAME_DTYPE = np.dtype([("pos", "<f4", 3), ("vel", "<f4", 3), ("center", "<f4", 3)], align=False)
assert FRAME_DTYPE.itemsize == 36


def lib_path() -> str:
    return os.environ.get("TRACEIT_B200_LIB", os.path.join(HERE, "libtraceit_b200.so"))


def build_extension(jobs: int = 4) -> str:
    """Compile libtraceit_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", os.path.join(HERE, "csrc"), f"-j{jobs}", "-s"], check=True)
    return os.path.join(HERE, "libtraceit_b200.so")


_lib = None


def lib():
    """The loaded C-ABI library.  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise TraceItError(2, f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(there is no CPU fallback)")
        L = C.CDLL(path)
        vp, u64p, u32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        L.tr_abi_version.restype = C.c_int
        L.tr_last_error.restype = C.c_char_p
        L.tr_device_count.argtypes = [C.POINTER(C.c_int)]
        L.tr_default_params.argtypes = [C.POINTER(Params)]
        L.tr_world_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(vp)]
        L.tr_comm_unique_id.argtypes = [vp]
        L.tr_world_create_sharded.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]
        L.tr_world_destroy.argtypes = [vp]
        L.tr_world_set_params.argtypes = [vp, C.POINTER(Params)]
        L.tr_world_get_params.argtypes = [vp, C.POINTER(Params)]
        L.tr_world_set_stream.argtypes = [vp, vp]
        L.tr_body_add.argtypes = [vp, vp, u32p]
        L.tr_bodies_add.argtypes = [vp, C.c_uint64, vp, u32p]
        L.tr_world_size.argtypes = [vp, u64p]
        L.tr_world_owned_range.argtypes = [vp, u64p, u64p]
        L.tr_to_spherical.argtypes = [C.c_float, C.c_float, C.c_float, vp]
        L.tr_area_from_size.argtypes = [C.c_float]
        L.tr_area_from_size.restype = C.c_float
        L.tr_step.argtypes = [vp, C.c_int]
        L.tr_sync.argtypes = [vp]
        L.tr_step_host.argtypes = [vp, C.c_uint64, vp, C.c_int]
        L.tr_step_frame.argtypes = [vp, C.c_uint64, vp, vp, C.c_int]
        L.tr_read_state.argtypes = [vp, C.c_uint64, C.c_uint64, vp, vp]
        L.tr_read_bodies.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
        L.tr_write_bodies.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
        L.tr_set_force.argtypes = [vp, C.c_uint32, C.c_float, C.c_float, C.c_float]
        L.tr_world_enable_trajectory.argtypes = [vp, C.c_uint32]
        L.tr_read_trajectory.argtypes = [vp, C.c_uint32, vp, C.c_uint64, u64p]
        L.tr_read_contacts.argtypes = [vp, vp, C.c_uint64, u64p]
        L.tr_debug_grid.argtypes = [vp, vp, vp, u64p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int32)]
        L.tr_debug_candidates.argtypes = [vp, vp, C.c_uint64, u64p]
        L.tr_debug_gravity.argtypes = [vp, vp]
        L.tr_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.tr_reset_stats.argtypes = [vp]
        L.tr_enable_timing.argtypes = [vp, C.c_int]
        L.tr_measure_fma_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        if L.tr_abi_version() != 1:
            raise TraceItError(1, "ABI version mismatch")
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise TraceItError(rc, lib().tr_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_params(**over) -> Params:
    p = Params()
    _check(lib().tr_default_params(C.byref(p)))
    for k, v in over.items():
        if k in ("g", "grid_origin"):
            arr = getattr(p, k)
            for i in range(3):
                arr[i] = float(v[i])
        else:
            setattr(p, k, v)
    return p


def new_bodies(n: int) -> np.ndarray:
    """Zeroed descriptors with the creation-form defaults (Cd 0.42, rest 0.5; reference
    src/main.cpp:1209,1306) and unit mass."""
    b = np.zeros(n, dtype=BODY_DTYPE)
    b["Cd"] = np.float32(0.42)
    b["rest"] = np.float32(0.5)
    b["mass"] = np.float32(1.0)
    return b


def to_spherical(elev_deg: float, azim_deg: float, magnitude: float) -> np.ndarray:
    out = np.zeros(3, dtype=np.float32)
    _check(lib().tr_to_spherical(elev_deg, azim_deg, magnitude, _ptr(out)))
    return out


def area_from_size(size: float) -> np.float32:
    return np.float32(lib().tr_area_from_size(size))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().tr_comm_unique_id(buf))
    return buf.raw


def measure_fma_peak(device: int = 0):
    ops, mhz = C.c_double(0), C.c_double(0)
    _check(lib().tr_measure_fma_peak(device, C.byref(ops), C.byref(mhz)))
    return ops.value, mhz.value


class World:
    """Opaque tr_world handle: the replacement for the reference's ``world w`` +
    ``updatePhysics(w, coords)`` (reference src/main.cpp:118-126, 748)."""

    def __init__(self, params: Params | None = None, device: int = 0, rank: int = 0, nranks: int = 1,
                 unique_id: bytes | None = None):
        self._L = lib()
        self._h = C.c_void_p()
        p = params if params is not None else default_params()
        if nranks > 1:
            assert unique_id is not None and len(unique_id) == 128
            _check(self._L.tr_world_create_sharded(C.byref(p), device, rank, nranks, unique_id, C.byref(self._h)))
        else:
            _check(self._L.tr_world_create(C.byref(p), device, C.byref(self._h)))

    # -- lifetime
    def close(self) -> None:
        if self._h:
            self._L.tr_world_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- parameters
    @property
    def params(self) -> Params:
        p = Params()
        _check(self._L.tr_world_get_params(self._h, C.byref(p)))
        return p

    @params.setter
    def params(self, p: Params) -> None:
        _check(self._L.tr_world_set_params(self._h, C.byref(p)))

    def set_stream(self, cuda_stream: int | None) -> None:
        _check(self._L.tr_world_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    # -- bodies
    def add(self, bodies: np.ndarray) -> int:
        assert bodies.dtype == BODY_DTYPE
        bodies = np.ascontiguousarray(bodies)
        first = C.c_uint32(0)
        _check(self._L.tr_bodies_add(self._h, len(bodies), _ptr(bodies), C.byref(first)))
        return first.value

    def __len__(self) -> int:
        n = C.c_uint64(0)
        _check(self._L.tr_world_size(self._h, C.byref(n)))
        return n.value

    def owned_range(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(self._L.tr_world_owned_range(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- stepping
    def step(self, nsteps: int = 1, sync: bool = True) -> None:
        _check(self._L.tr_step(self._h, nsteps))
        if sync:
            _check(self._L.tr_sync(self._h))

    def sync(self) -> None:
        _check(self._L.tr_sync(self._h))

    def step_host(self, bodies: np.ndarray, nsteps: int = 1) -> np.ndarray:
        assert bodies.dtype == BODY_DTYPE and bodies.flags.c_contiguous
        _check(self._L.tr_step_host(self._h, len(bodies), _ptr(bodies), nsteps))
        return bodies

    def step_frame(self, out: np.ndarray, forces: np.ndarray | None = None, nsteps: int = 1) -> np.ndarray:
        """tr_step_frame: forces (n,3) float32 up (optional), nsteps, pos/vel/centre down into `out` (FRAME_DTYPE)."""
        assert out.dtype == FRAME_DTYPE and out.flags.c_contiguous
        if forces is not None:
            assert forces.dtype == np.float32 and forces.flags.c_contiguous and forces.shape == (len(out), 3)
        _check(self._L.tr_step_frame(self._h, len(out), _ptr(forces), _ptr(out), nsteps))
        return out

    # -- state
    def read_state(self, first: int = 0, count: int | None = None, vel: bool = True):
        if count is None:
            count = len(self) - first
        pos = np.empty((count, 3), dtype=np.float32)
        v = np.empty((count, 3), dtype=np.float32) if vel else None
        _check(self._L.tr_read_state(self._h, first, count, _ptr(pos), _ptr(v)))
        return pos, v

    def read_bodies(self, first: int = 0, count: int | None = None) -> np.ndarray:
        if count is None:
            count = len(self) - first
        out = np.zeros(count, dtype=BODY_DTYPE)
        _check(self._L.tr_read_bodies(self._h, first, count, _ptr(out)))
        return out

    def write_bodies(self, bodies: np.ndarray, first: int = 0) -> None:
        assert bodies.dtype == BODY_DTYPE
        bodies = np.ascontiguousarray(bodies)
        _check(self._L.tr_write_bodies(self._h, first, len(bodies), _ptr(bodies)))

    def set_force(self, idx: int, f) -> None:
        _check(self._L.tr_set_force(self._h, idx, float(f[0]), float(f[1]), float(f[2])))

    # -- trajectory history (object::trajectory)
    def enable_trajectory(self, max_points: int = 300) -> None:
        _check(self._L.tr_world_enable_trajectory(self._h, max_points))

    def trajectory(self, idx: int) -> np.ndarray:
        n = C.c_uint64(0)
        _check(self._L.tr_read_trajectory(self._h, idx, None, 0, C.byref(n)))
        out = np.empty((n.value, 3), dtype=np.float32)
        if n.value:
            _check(self._L.tr_read_trajectory(self._h, idx, _ptr(out), n.value, C.byref(n)))
        return out

    # -- parity / observability
    def contacts(self) -> np.ndarray:
        n = C.c_uint64(0)
        _check(self._L.tr_read_contacts(self._h, None, 0, C.byref(n)))
        out = np.empty((n.value, 2), dtype=np.int32)
        if n.value:
            _check(self._L.tr_read_contacts(self._h, _ptr(out), n.value, C.byref(n)))
        return out

    def debug_grid(self):
        n = len(self)
        keys = np.empty(n, dtype=np.uint32)
        ids = np.empty(n, dtype=np.uint32)
        m = C.c_uint64(0)
        cell, inv, bits = C.c_float(0), C.c_float(0), C.c_int32(0)
        _check(self._L.tr_debug_grid(self._h, _ptr(keys), _ptr(ids), C.byref(m), C.byref(cell), C.byref(inv), C.byref(bits)))
        return {"keys": keys, "sorted_ids": ids[: m.value].copy(), "cell": cell.value, "inv_cell": inv.value,
                "bits": bits.value}

    def debug_candidates(self, want_pairs: bool = True):
        n = C.c_uint64(0)
        _check(self._L.tr_debug_candidates(self._h, None, 0, C.byref(n)))
        if not want_pairs:
            return n.value, None
        out = np.empty((n.value, 2), dtype=np.int32)
        if n.value:
            _check(self._L.tr_debug_candidates(self._h, _ptr(out), n.value, C.byref(n)))
        return n.value, out

    def debug_gravity(self) -> np.ndarray:
        _, cnt = self.owned_range()
        out = np.empty((cnt, 3), dtype=np.float32)
        _check(self._L.tr_debug_gravity(self._h, _ptr(out)))
        return out

    def stats(self) -> Stats:
        s = Stats()
        _check(self._L.tr_get_stats(self._h, C.byref(s)))
        return s

    def reset_stats(self) -> None:
        _check(self._L.tr_reset_stats(self._h))

    def enable_timing(self, on: bool = True) -> None:
        _check(self._L.tr_enable_timing(self._h, 1 if on else 0))
