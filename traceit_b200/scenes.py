"""Synthetic scenes for the BASELINE.json configurations.

All generators are deterministic functions of their arguments (numpy ``RandomState`` =
MT19937, the same engine family the survey names) and return ``tr_body_desc`` arrays.
They are host-side setup, shared by the tests and ``bench.py``.
"""
from __future__ import annotations

import ctypes
import ctypes.util

import numpy as np

from . import BODY_DTYPE, new_bodies

SEED_CFG2 = 0xC0FFEE02
SEED_CFG3 = 0xC0FFEE03
SEED_CFG4 = 0xC0FFEE04
SEED_CFG5 = 0xC0FFEE05

PI_F32_AREA_HALF = np.float32(np.pi * 0.5 * 0.5)  # area = M_PI*size*size with size 0.5 (double math, narrowed)


_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.cosf.restype = _libm.sinf.restype = ctypes.c_float
_libm.cosf.argtypes = _libm.sinf.argtypes = [ctypes.c_float]


def _to_spherical(elev, azim, magn):
    """toSpherical, reference src/main.cpp:1080-1089, restated for scene setup without needing
    the CUDA library: the factor (M_PI/180) is a double, the product is narrowed to float,
    trig is libm's cosf/sinf (numpy's float32 trig differs by an ulp), float products left to
    right.  tr_to_spherical in the C ABI is the product version of the same helper."""
    rad_e = np.float32(np.float64(np.float32(elev)) * (np.pi / 180))
    rad_a = np.float32(np.float64(np.float32(azim)) * (np.pi / 180))
    m = np.float32(magn)
    ce, se = np.float32(_libm.cosf(float(rad_e))), np.float32(_libm.sinf(float(rad_e)))
    ca, sa = np.float32(_libm.cosf(float(rad_a))), np.float32(_libm.sinf(float(rad_a)))
    return np.array([m * ce * ca, m * se, m * ce * sa], dtype=np.float32)


def default_scene() -> np.ndarray:
    """Config 1: the reference's seeded world plus two UI-style projectiles.

    earth exactly as reference src/main.cpp:1482-1491; projectiles within the creation-form
    clamps (src/main.cpp:1216-1224) with the collider the form builds (src/main.cpp:1302-1307:
    AABB, size {s,s,s}, rest 0.5).  velocities use the C helper tr_to_spherical when the
    library is loaded by the caller; here the same arithmetic is restated in numpy.
    """
    b = new_bodies(3)
    # earth ("planet")
    b[0]["pos"] = (0, -20, 0)
    b[0]["mass"] = 1000
    b[0]["Cd"] = 1000.0
    b[0]["area"] = 1000.0
    b[0]["stationary"] = 1
    b[0]["center"] = (0, -20, 0)
    b[0]["size"] = (10000, 5, 10000)
    b[0]["isSphere"] = 0
    b[0]["isStatic"] = 0
    b[0]["rest"] = 1.0
    # P1: mass 10, elev 45, azim 0, magnitude 30, size 1, at the origin
    b[1]["mass"] = 10
    b[1]["vel"] = _to_spherical(45, 0, 30)
    b[1]["area"] = np.float32(np.pi * 1.0 * 1.0)
    b[1]["size"] = (1, 1, 1)
    # P2: mass 5, elev 60, azim 90, magnitude 20, size 2, at (5,0,0)
    b[2]["mass"] = 5
    b[2]["pos"] = (5, 0, 0)
    b[2]["center"] = (5, 0, 0)
    b[2]["vel"] = _to_spherical(60, 90, 20)
    b[2]["area"] = np.float32(np.pi * 2.0 * 2.0)
    b[2]["size"] = (2, 2, 2)
    return b


def random_cloud(n: int, seed: int, spheres: bool, mass_scale: float = 1.0, radius: float = 0.5,
                 density_l: float = 4.0) -> np.ndarray:
    """Configs 2-4: uniform cloud in a cube of half-width L = density_l * cbrt(n); velocities
    uniform in [-1,1]^3; masses uniform in [1,10]*mass_scale; Cd 0.42; size 0.5."""
    rs = np.random.RandomState(seed & 0xFFFFFFFF)
    L = density_l * np.cbrt(float(n))
    b = new_bodies(n)
    b["pos"] = rs.uniform(-L, L, size=(n, 3)).astype(np.float32)
    b["vel"] = rs.uniform(-1.0, 1.0, size=(n, 3)).astype(np.float32)
    b["mass"] = (rs.uniform(1.0, 10.0, size=n) * mass_scale).astype(np.float32)
    b["area"] = PI_F32_AREA_HALF
    b["center"] = b["pos"]
    b["size"] = np.float32(radius)
    b["radius"] = np.float32(radius)
    b["isSphere"] = 1 if spheres else 0
    return b


def cfg2_cloud(n: int = 65536) -> np.ndarray:
    """64K-body random cloud, collisions off; masses [1,10] with G = 1 (see params_for)."""
    return random_cloud(n, SEED_CFG2, spheres=True)


def cfg3_cloud(n: int = 1 << 20) -> np.ndarray:
    """1M-body cloud with radius-0.5 spheres (sparse contacts)."""
    return random_cloud(n, SEED_CFG3, spheres=True)


def cfg4_cloud(n: int = 1 << 22) -> np.ndarray:
    """4M-body cloud for the sharded gravity runs."""
    return random_cloud(n, SEED_CFG4, spheres=True)


def dense_lattice(n_side, seed: int = SEED_CFG5, pitch: float = 0.98, jitter: float = 0.01,
                  shuffle_ids: bool = False) -> np.ndarray:
    """Config 5: jittered lattice of radius-0.5 spheres at pitch 0.98 (every body overlaps its
    ~6 face neighbours), velocities uniform in [-0.1,0.1]^3.  n_side is an int (cube) or a
    (nx, ny, nz) tuple; ids run x-fastest in lattice order unless shuffled."""
    dims = (n_side,) * 3 if np.isscalar(n_side) else tuple(int(d) for d in n_side)
    nx, ny, nz = dims
    rs = np.random.RandomState(seed & 0xFFFFFFFF)
    n = nx * ny * nz
    idx = np.arange(n)
    ix, iy, iz = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    centre = np.array([(nx - 1) / 2.0, (ny - 1) / 2.0, (nz - 1) / 2.0])
    base = (np.stack([ix, iy, iz], axis=1).astype(np.float64) - centre) * pitch
    pos = (base + rs.uniform(-jitter, jitter, size=(n, 3))).astype(np.float32)
    vel = rs.uniform(-0.1, 0.1, size=(n, 3)).astype(np.float32)
    if shuffle_ids:
        perm = rs.permutation(n)
        pos, vel = pos[perm], vel[perm]
    b = new_bodies(n)
    b["pos"] = pos
    b["vel"] = vel
    b["mass"] = np.float32(1.0)
    b["area"] = PI_F32_AREA_HALF
    b["center"] = pos
    b["size"] = np.float32(0.5)
    b["radius"] = np.float32(0.5)
    b["isSphere"] = 1
    return b


def cfg5_dense(n: int = 1 << 20) -> np.ndarray:
    """1,048,576 = 128 x 128 x 64 packed spheres (smaller powers of two scale the box down)."""
    e = int(np.log2(n))
    assert 1 << e == n, "cfg5 sizes are powers of two"
    ex = (e + 2) // 3
    ey = (e - ex + 1) // 2
    ez = e - ex - ey
    return dense_lattice((1 << ex, 1 << ey, 1 << ez))


assert new_bodies(0).dtype == BODY_DTYPE
