"""CPU tests of the oracle itself: the restated C step against (a) the committed golden
vectors, which were produced by the verbatim reference build, and (b) the verbatim
reference library directly when it is present (oracle/_ref, built where /root/reference
exists).  This is what pins the oracle."""
import os
import sys

import numpy as np
import pytest

from conftest import assert_bits_equal, load_golden, mixed_cloud
from traceit_b200 import scenes

FIELDS = ("pos", "vel", "force", "center")


def _check_state(po, got, want, what):
    for f in FIELDS:
        assert_bits_equal(got[f], want[f], f"{what}:{f}")
    assert po.state_hash(got) == po.state_hash(want)


def test_scene_generators_match_goldens():
    g = load_golden("default_scene.npz")
    assert scenes.default_scene().tobytes() == g["initial"].tobytes()
    g = load_golden("sphere_cloud_2048_s1.npz")
    assert scenes.random_cloud(2048, 1234, spheres=True, density_l=0.6).tobytes() == g["initial"].tobytes()
    g = load_golden("mixed_cloud_1024_s40.npz")
    assert mixed_cloud().tobytes() == g["initial"].tobytes()
    g = load_golden("dense_lattice_12_s20.npz")
    assert scenes.dense_lattice(12).tobytes() == g["initial"].tobytes()


def test_default_scene_golden(oracle):
    """Config 1: 1000 steps of the reference's default-style scene, bit for bit."""
    g = load_golden("default_scene.npz")
    p = oracle.default_params()
    b = g["initial"].copy()
    done = 0
    for s, want in zip(g["trace_steps"], g["trace_p1"]):
        oracle.step(b, p, int(s) - done)
        done = int(s)
        assert_bits_equal(b[1]["pos"], want["pos"], f"P1 pos @ step {s}")
        assert_bits_equal(b[1]["vel"], want["vel"], f"P1 vel @ step {s}")
    oracle.step(b, p, 1000 - done)
    _check_state(oracle, b, g["final"], "default scene")
    assert oracle.state_hash(b) == int(g["hash"])
    # the reference's inverted approach test: the projectile falls through the slab
    assert b[1]["pos"][1] < -100 and b[0]["vel"][1] > 7


@pytest.mark.parametrize("name,steps", [("sphere_cloud_2048_s1.npz", 1), ("sphere_cloud_2048_s50.npz", 50),
                                        ("mixed_cloud_1024_s40.npz", 40), ("dense_lattice_12_s20.npz", 20)])
def test_cloud_goldens(oracle, name, steps):
    g = load_golden(name)
    b = g["initial"].copy()
    oracle.step(b, oracle.default_params(), steps)
    _check_state(oracle, b, g["final"], name)
    assert np.array_equal(oracle.contacts(b), g["contacts"])
    assert oracle.state_hash(b) == int(g["hash"])


def test_resolve_list_equals_sweep(oracle):
    """Replaying the precomputed contact list is bit-identical to the O(N^2) sweep."""
    g = load_golden("dense_lattice_12_s20.npz")
    p = oracle.default_params()
    a = g["initial"].copy()
    b = g["initial"].copy()
    for _ in range(5):
        oracle.integrate(a, p)
        oracle.resolve(a)
        oracle.integrate(b, p)
        oracle.resolve_list(b, oracle.contacts(b))
    _check_state(oracle, a, b, "resolve_list")


def test_restated_equals_verbatim_reference(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libtraceit_ref.so not built (needs /root/reference)")
    rs = np.random.RandomState(5)
    for n, dens, steps in ((300, 0.5, 30), (1500, 0.7, 10)):
        b = scenes.random_cloud(n, int(rs.randint(1 << 30)), spheres=True, density_l=dens)
        b["isSphere"][rs.rand(n) < 0.2] = 0
        b["isStatic"][rs.rand(n) < 0.1] = 1
        b["stationary"][rs.rand(n) < 0.1] = 1
        b["force"] = rs.uniform(-5, 5, size=(n, 3)).astype(np.float32)
        rw = oracle.RefWorld(0.7)
        rw.add(b)
        rw.step(steps)
        want = rw.read()
        got = b.copy()
        oracle.step(got, oracle.default_params(atm_density=np.float32(0.7)), steps)
        _check_state(oracle, got, want, f"n={n}")
        assert np.array_equal(rw.contacts(), oracle.contacts(got))
        rw.close()


def test_reference_o0_equals_o2(oracle):
    if not (oracle.ref_available() and oracle.ref_available(o0=True)):
        pytest.skip("reference builds missing")
    b = scenes.random_cloud(512, 9, spheres=True, density_l=0.6)
    out = []
    for o0 in (False, True):
        rw = oracle.RefWorld(1.2, o0=o0)
        rw.add(b)
        rw.step(20)
        out.append(rw.read())
        rw.close()
    assert oracle.state_hash(out[0]) == oracle.state_hash(out[1])


def test_creation_helpers_golden(oracle):
    g = load_golden("creation_helpers.npz")
    for a, want in zip(g["angles"], g["spherical"]):
        assert_bits_equal(scenes._to_spherical(*a), want, f"toSpherical{tuple(a)}")
    for s, want in zip(g["sizes"], g["areas"]):
        assert_bits_equal(np.float32(np.pi * float(s) * float(s)), want, f"area({s})")


def test_pair_gravity_f32_vs_f64(oracle):
    """The fp32 restatement of the dead pair-gravity block against the fp64 yardstick."""
    b = scenes.random_cloud(2000, 3, spheres=True)
    f32 = oracle.pair_gravity_f32(b, 1.0).astype(np.float64)
    f64 = oracle.pair_gravity_f64(b, 1.0)
    rel = np.linalg.norm(f32 - f64) / np.linalg.norm(f64)
    assert rel < 1e-5, rel
    # Newton's third law in the fp64 sum
    assert np.abs(f64.sum(axis=0)).max() < 1e-9 * np.abs(f64).sum()


def test_pair_gravity_coincident_bodies_are_nan(oracle):
    """d = 0 between DISTINCT bodies: the reference formula gives G*m*m/0 * normalize(0) = NaN."""
    b = scenes.random_cloud(4, 3, spheres=True)
    b["pos"][1] = b["pos"][0]
    f = oracle.pair_gravity_f32(b, 1.0)
    assert np.isnan(f[0]).all() and np.isnan(f[1]).all() and np.isfinite(f[2:]).all()


def test_grid_contacts_equal_brute_force(oracle):
    """CPU grid restatement finds exactly the brute-force contact set (incl. wrap-around)."""
    # bits code: b = the same on every axis, or bx | by << 8 | bz << 16
    for n, dens, bits in ((3000, 0.5, 4), (3000, 0.5, 2), (5000, 2.0, 3), (4000, 0.7, 5 | (4 << 8) | (4 << 16)),
                          (3000, 0.5, 3 | (2 << 8) | (2 << 16))):
        b = scenes.random_cloud(n, 11 + (bits & 0xFF), spheres=True, density_l=dens)
        b["radius"] = np.random.RandomState(bits & 0xFF).uniform(0.1, 0.5, n).astype(np.float32)
        rmax = float(b["radius"].max())
        cell = np.float32(np.float32(2.0 * rmax) * np.float32(1.0078125))
        inv = np.float32(1.0) / cell
        cand, cont, nc, nk = oracle.grid_pairs(b, np.ones(n, np.uint8), (0, 0, 0), inv, bits)
        brute = oracle.contacts(b)
        assert np.array_equal(cont, brute), (n, bits)
        assert nc >= nk and nk == len(brute)
        # candidates are unique
        assert len(np.unique(cand, axis=0)) == len(cand)


def test_grid_sort_is_key_then_id(oracle):
    b = scenes.random_cloud(4096, 21, spheres=True)
    keys = oracle.grid_keys(b, (0, 0, 0), np.float32(0.99), 5)
    order, skeys = oracle.grid_sort(keys, np.arange(4096, dtype=np.int32), 5)
    ref = np.lexsort((np.arange(4096), keys))
    assert np.array_equal(order, ref.astype(np.int32))
    assert np.array_equal(skeys, keys[ref])


@pytest.mark.parametrize("name,n,dens", [("grid_2000_b5", 2000, 0.6), ("grid_3000_b544", 3000, 0.7)])
def test_grid_definition_goldens(oracle, name, n, dens):
    """The grid has no reference counterpart, so its DEFINITION (keys, stable order, candidate set) is
    pinned by committed vectors (tests/golden/make_grid_golden.py, whose contact sets were checked
    against the verbatim reference's sweep); one case has the same bits on every axis, one has not."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_grid_golden as mg
    g = load_golden(name + ".npz")
    b = mg.scene(n, dens)
    bits, inv = int(g["bits"]), np.float32(g["inv_cell"])
    keys = oracle.grid_keys(b, (0, 0, 0), inv, bits)
    assert np.array_equal(keys, g["keys"])
    order, _ = oracle.grid_sort(keys, np.arange(n, dtype=np.int32), bits)
    assert np.array_equal(order, g["order"])
    cand, cont, nc, _ = oracle.grid_pairs(b, np.ones(n, np.uint8), (0, 0, 0), inv, bits)
    assert nc == int(g["n_candidates"]) and mg.fnv1a(cand) == int(g["candidates_fnv1a"])
    assert np.array_equal(cont, g["contacts"])


@pytest.mark.parametrize("name,seed", [("traj_cfg2_64k_k100.npz", "SEED_CFG2"), ("traj_cfg3cut_64k_k20.npz", "SEED_CFG3")])
def test_trajectory_goldens_are_consistent(oracle, name, seed):
    """The 64K long-horizon goldens (tests/golden/make_traj_golden.py, minutes of oracle time) are only
    REPLAYED on the GPU box; here: their inputs are what the scene generator makes today, and their
    shapes are sane.  One oracle step from the same inputs is also checked against a sampled body's
    displacement bound (a cheap guard against a golden generated with other parameters)."""
    g = load_golden(name)
    b = scenes.random_cloud(65536, getattr(scenes, seed), spheres=True)
    assert oracle.state_hash(b) == int(g["input_hash"])
    ids = g["ids"]
    assert np.all(np.diff(ids) > 0) and ids.max() < 65536
    assert g["pos"].shape == (len(ids), 3) and g["vel"].shape == (len(ids), 3)
    assert np.isfinite(g["pos"]).all() and np.isfinite(g["vel"]).all()
    # nothing travels further than K * dt * (a generous speed bound: initial speeds are <= sqrt(3); the reference's
    # inverted approach test pumps energy into bodies in contact, seen up to ~40 here)
    k = int(g["steps"])
    vmax = 64.0 if "offsets" in g.files else 8.0
    assert np.abs(g["pos"] - b["pos"][ids]).max() < k * 0.01666 * vmax
    if "offsets" in g.files:
        offs = g["offsets"]
        assert len(offs) == k + 1 and offs[0] == 0 and offs[-1] == len(g["contacts"]) and np.all(np.diff(offs) >= 0)
        c = g["contacts"]
        assert np.all(c[:, 0] < c[:, 1])
        # the first step's contact set is cheap to recompute: integrate once, brute-force detection
        w = b.copy()
        oracle.step(w, oracle.default_params(flags=oracle.PAIR_GRAVITY, G=np.float32(g["G"]), g=(0, 0, 0)), 1)
        assert np.array_equal(oracle.contacts(w), c[offs[0]:offs[1]])
