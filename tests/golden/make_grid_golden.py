"""Golden vectors of the uniform-grid DEFINITION (cell keys, stable (key,id) order, candidate pairs).

The reference has no grid (brute force, src/main.cpp:674-675), so these come from the restated oracle
(oracle/restated.c orc_grid_*); what ties them to the reference is the contact set, which this script
checks against the VERBATIM reference's own sweep (oracle/_ref) before writing anything.

    make -C oracle ref && python tests/golden/make_grid_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from traceit_b200 import scenes  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = (("grid_2000_b5", 2000, 0.6, 5), ("grid_3000_b544", 3000, 0.7, 5 | (4 << 8) | (4 << 16)))


def scene(n, dens):
    b = scenes.random_cloud(n, 4242 + n, spheres=True, density_l=dens)
    b["radius"] = np.random.RandomState(n).uniform(0.2, 0.5, n).astype(np.float32)
    return b


def fnv1a(a: np.ndarray) -> int:
    h = 0xCBF29CE484222325
    for byte in np.ascontiguousarray(a).view(np.uint8).ravel():
        h = ((h ^ int(byte)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


if __name__ == "__main__":
    for name, n, dens, bits in CASES:
        b = scene(n, dens)
        cell = np.float32(np.float32(2.0 * float(b["radius"].max())) * np.float32(1.0078125))
        inv = np.float32(1.0) / cell
        keys = po.grid_keys(b, (0, 0, 0), inv, bits)
        order, _ = po.grid_sort(keys, np.arange(n, dtype=np.int32), bits)
        cand, cont, nc, nk = po.grid_pairs(b, np.ones(n, np.uint8), (0, 0, 0), inv, bits)
        # the verbatim reference's own O(N^2) sweep finds the same contacts
        rw = po.RefWorld(1.2)
        rw.add(b)
        ref_contacts = rw.contacts()             # the reference's own predicates over all i<j
        rw.close()
        assert np.array_equal(cont, ref_contacts), name
        assert np.array_equal(cont, po.contacts(b)), name
        np.savez_compressed(os.path.join(OUT, name + ".npz"), bits=np.int64(bits), inv_cell=inv, keys=keys, order=order,
                            n_candidates=np.int64(nc), candidates_fnv1a=np.uint64(fnv1a(cand)), contacts=cont)
        print(name, "candidates", nc, "contacts", nk)
