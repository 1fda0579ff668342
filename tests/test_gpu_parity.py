"""GPU parity tests (run on a B200 with `-m gpu`): every call goes through the C ABI of
libtraceit_b200.so and is compared with the CPU oracle on the same seeded inputs.

Bars (SURVEY.md 8d / north_star):
  * integrator, collision detection, contact set, collision response: BIT-EXACT;
  * cell keys, sort order, candidate-pair set: BIT-EXACT against the CPU grid restatement;
  * pair gravity (dead code in the reference, restated): force rel. L2 <= 1e-5 against the
    fp64 restatement; positions/velocities after K steps: max error <= 1e-4 relative to the
    cloud extent / rms speed against the fp32 restated oracle.
"""
import numpy as np
import pytest

from conftest import assert_bits_equal, load_golden, mixed_cloud
from traceit_b200 import scenes

pytestmark = pytest.mark.gpu

FIELDS = ("pos", "vel", "force", "center")
FORCE_REL_L2_TOL = 1e-5      # GPU pair-gravity force vs fp64 restatement
TRAJ_REL_TOL = 1e-4          # pos/vel after K steps vs fp32 restated oracle


def gpu_run(tr, bodies, steps, **pover):
    p = tr.default_params(**pover)
    with tr.World(p) as w:
        w.add(bodies)
        w.step(steps)
        out = w.read_bodies()
        contacts = w.contacts() if (p.flags & tr.COLLISIONS) else None
        stats = w.stats()
    return out, contacts, stats


def check_state(got, want, what):
    for f in FIELDS:
        assert_bits_equal(got[f], want[f], f"{what}:{f}")


# ---------------------------------------------------------------- O(N) integrator
def test_integrator_bit_exact(tr, oracle):
    rs = np.random.RandomState(42)
    n = 10007
    b = scenes.random_cloud(n, 5, spheres=True)
    b["stationary"][rs.rand(n) < 0.1] = 1
    b["force"] = rs.uniform(-20, 20, size=(n, 3)).astype(np.float32)
    b["vel"][rs.rand(n) < 0.05] = 0          # exercises the speedSq <= 1e-4 branch
    b["vel"][7] = (0.005, 0.0, 0.0)
    b["Cd"] = rs.uniform(0, 2, n).astype(np.float32)
    got, _, st = gpu_run(tr, b, 25, flags=0, atm_density=np.float32(0.9))
    want = b.copy()
    oracle.step(want, oracle.default_params(flags=0, atm_density=np.float32(0.9)), 25)
    check_state(got, want, "integrator")
    assert st.kernel_launches >= 25


def test_step_host_round_trip_equals_step(tr, oracle):
    b = scenes.random_cloud(3000, 8, spheres=True, density_l=0.6)
    want = b.copy()
    oracle.step(want, oracle.default_params(), 3)
    p = tr.default_params()
    with tr.World(p) as w:
        cur = b.copy()
        for _ in range(3):
            w.step_host(cur, 1)      # upload AoS -> step -> read back, the shim's per-frame path
    check_state(cur, want, "step_host")


# ---------------------------------------------------------------- goldens from the verbatim reference
def test_default_scene_golden_gpu(tr):
    """Config 1 on the GPU: the two box projectiles live in the grid (AABB predicate), the
    10000x5x10000 slab is oversize and sits in the side list; 1000 steps."""
    g = load_golden("default_scene.npz")
    # through the general pipeline (a 3-body world normally takes the single-CTA path, tested below)
    for extra in (0, tr.DEBUG_NO_SPARSE):
        got, contacts, st = gpu_run(tr, g["initial"].copy(), 1000, debug_flags=tr.DEBUG_NO_TINY | extra)
        check_state(got, g["final"], "default scene")
        assert np.array_equal(contacts, g["contacts"])
        assert st.side_list_size == 1


def test_default_scene_golden_single_cta_path(tr):
    """The same 1000 steps the way a 3-body world really runs: whole steps inside ONE kernel of one CTA
    (tiny.cuh), here as one launch of 1000 steps and as 1000 launches of one."""
    g = load_golden("default_scene.npz")
    got, contacts, st = gpu_run(tr, g["initial"].copy(), 1000)
    check_state(got, g["final"], "default scene, tr_step(w, 1000)")
    assert np.array_equal(contacts, g["contacts"])
    assert st.kernel_launches <= 3 and st.steps == 1000
    with tr.World(tr.default_params()) as w:
        w.add(g["initial"].copy())
        for _ in range(1000):
            w.step(1, sync=False)
        check_state(w.read_bodies(), g["final"], "default scene, 1000 x tr_step(w, 1)")
        assert np.array_equal(w.contacts(), g["contacts"])


@pytest.mark.parametrize("name,steps", [("sphere_cloud_2048_s1.npz", 1), ("sphere_cloud_2048_s50.npz", 50),
                                        ("mixed_cloud_1024_s40.npz", 40), ("dense_lattice_12_s20.npz", 20)])
@pytest.mark.parametrize("sparse", [True, False])
def test_cloud_goldens_gpu(tr, name, steps, sparse):
    """These steps have a few hundred contacts: replayed by the epilogue kernel's single CTA in shared memory
    (sparse=True, the default) or pushed through the grid-wide ordering + dataflow executor."""
    g = load_golden(name)
    got, contacts, st = gpu_run(tr, g["initial"].copy(), steps, debug_flags=0 if sparse else tr.DEBUG_NO_SPARSE)
    check_state(got, g["final"], name)
    assert np.array_equal(contacts, g["contacts"]), name


# ---------------------------------------------------------------- broad phase: keys, order, candidates
@pytest.mark.parametrize("n,dens,bits", [(20000, 0.7, 0), (20000, 0.7, 3), (5000, 3.0, 2), (4096, 0.5, 6),
                                         (20000, 0.7, 8)])  # bits 8: 16M cells, the streamed-span variant of the fused scan
def test_grid_keys_sort_candidates_bit_exact(tr, oracle, n, dens, bits):
    b = scenes.random_cloud(n, 100 + bits, spheres=True, density_l=dens)
    b["radius"] = np.random.RandomState(bits).uniform(0.2, 0.5, n).astype(np.float32)
    b["mass"][::17] = 0          # dropped from the grid
    b["isSphere"][::29] = 0      # boxes (size 0.5): share the grid, candidates are shape independent
    p = tr.default_params(flags=tr.COLLISIONS, grid_bits=bits)
    with tr.World(p) as w:
        w.add(b)
        g = w.debug_grid()
        ncand, cand = w.debug_candidates()
    gb = g["bits"]
    in_grid = ~(b["mass"] <= 0)
    keys = oracle.grid_keys(b, (0, 0, 0), g["inv_cell"], gb)
    assert np.array_equal(g["keys"][in_grid], keys[in_grid])
    assert (g["keys"][~in_grid] == 0xFFFFFFFF).all()
    order, _ = oracle.grid_sort(keys, np.nonzero(in_grid)[0].astype(np.int32), gb)
    assert np.array_equal(g["sorted_ids"].astype(np.int32), order)
    ocand, _, nc, _ = oracle.grid_pairs(b, in_grid.astype(np.uint8), (0, 0, 0), g["inv_cell"], gb, want_contacts=False)
    assert ncand == nc
    assert np.array_equal(cand, ocand)
    # the grid actually used is the documented one: extent = max(|radius| if sphere, |size|) in a mixed world
    rmax = max(float(b["radius"][in_grid & (b["isSphere"] != 0)].max()), float(np.abs(b["size"][in_grid]).max()))
    assert g["cell"] == np.float32(np.float32(2.0 * rmax) * np.float32(1.0078125))


@pytest.mark.parametrize("name,n,dens,grid_bits", [("grid_2000_b5", 2000, 0.6, 5), ("grid_3000_b544", 3000, 0.7, 0)])
def test_grid_definition_goldens_gpu(tr, name, n, dens, grid_bits):
    """The committed grid vectors (tests/golden/make_grid_golden.py): same bits on every axis when
    forced, (5,4,4) chosen automatically for 3000 bodies."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_grid_golden as mg
    g = load_golden(name + ".npz")
    b = mg.scene(n, dens)
    with tr.World(tr.default_params(flags=tr.COLLISIONS, grid_bits=grid_bits)) as w:
        w.add(b)
        gg = w.debug_grid()
        ncand, cand = w.debug_candidates()
    assert gg["bits"] == int(g["bits"]) and np.float32(gg["inv_cell"]) == np.float32(g["inv_cell"])
    assert np.array_equal(gg["keys"], g["keys"])
    assert np.array_equal(gg["sorted_ids"].astype(np.int32), g["order"])
    assert ncand == int(g["n_candidates"])
    assert mg.fnv1a(np.ascontiguousarray(cand, dtype=np.int32)) == int(g["candidates_fnv1a"])


@pytest.mark.parametrize("n,dens", [(30000, 0.6), (30000, 4.0)])
def test_contact_set_equals_brute_force(tr, oracle, n, dens):
    b = scenes.random_cloud(n, 321, spheres=True, density_l=dens)
    b["isStatic"][::5] = 1
    b["isSphere"][::1000] = 0
    b["size"][::1000] = (3.0, 0.2, 3.0)
    b["radius"][5] = 40.0          # oversize sphere -> side list
    want = b.copy()
    oracle.integrate(want, oracle.default_params())
    brute = oracle.contacts(want)
    p = tr.default_params()
    with tr.World(p) as w:
        w.add(b)
        w.step(1)
        got = w.contacts()
        st = w.stats()
    assert np.array_equal(got, brute)
    assert st.side_list_size == 1          # only the oversize sphere; the 3.0 x 0.2 x 3.0 boxes share the grid
    assert st.last_contacts == len(brute)


def test_out_of_domain_bodies_use_side_list(tr, oracle):
    b = scenes.random_cloud(4000, 77, spheres=True, density_l=0.6)
    b["pos"][10] = (1e6, 0, 0)
    b["pos"][11] = (1e6 + 0.25, 0, 0)   # touches body 10, far outside the grid domain
    b["pos"][12] = (np.nan, 0, 0)
    b["center"] = b["pos"]
    want = b.copy()
    oracle.step(want, oracle.default_params(), 2)
    got, contacts, st = gpu_run(tr, b, 2)
    assert st.side_list_size >= 2      # the two far bodies; the NaN body is 'lost' (can never touch anything finite)
    for f in FIELDS:
        assert_bits_equal(got[f], want[f], f"domain:{f}")
    assert np.array_equal(contacts, oracle.contacts(want))


def test_far_bodies_meet_each_other_and_the_domain_boundary(tr, oracle):
    """Bodies outside the grid domain (|centre - origin| / cell >= 32768) are tested against each
    other and against the grid bodies next to the domain boundary only; the contact set and the
    response must still equal the reference's O(N^2) sweep."""
    b = scenes.random_cloud(3000, 78, spheres=True, density_l=0.6)
    cell = np.float32(np.float32(2.0 * 0.5) * np.float32(1.0078125))
    edge = float(cell) * 32768.0
    k = 100
    for axis, sign in ((0, 1), (1, -1), (2, 1)):
        for d_in, d_out in ((0.45, 0.40), (0.2, 0.7), (1.6, 0.1)):   # touching, touching, apart
            p_in = np.zeros(3, np.float32); p_out = np.zeros(3, np.float32)
            p_in[axis] = sign * (edge - d_in); p_out[axis] = sign * (edge + d_out)
            b["pos"][k] = p_in; b["pos"][k + 1] = p_out
            k += 2
    # a far cluster: chain of touching spheres, one static pair, one box
    for q in range(12):
        b["pos"][k + q] = (2.0e5 + 0.75 * q, -3.0e5, 0.5 * (q % 2))
    b["isStatic"][k + 3] = 1; b["isStatic"][k + 4] = 1
    b["isSphere"][k + 7] = 0
    b["center"] = b["pos"]
    for steps in (1, 3):
        want = b.copy()
        oracle.step(want, oracle.default_params(), steps)
        got, contacts, st = gpu_run(tr, b.copy(), steps)
        assert st.side_list_size == 9 + 12 if steps == 1 else st.side_list_size >= 12   # the response moves boundary bodies in and out
        oc = oracle.contacts(want)
        assert np.array_equal(contacts, oc)
        if steps == 1:   # both kinds of pair are really there
            assert any(100 <= i < 118 and 100 <= j < 118 for i, j in oc), "boundary pairs"
            assert any(118 <= i < 130 and 118 <= j < 130 for i, j in oc), "far-far pairs"
        for f in FIELDS:
            assert_bits_equal(got[f], want[f], f"far:{f}")


# ---------------------------------------------------------------- response order
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_dense_lattice_response_bit_exact(tr, oracle, mode):
    """Config 5 (small): ~3 contacts per body, lattice-ordered ids = long Gauss-Seidel chains.
    mode 0 = dataflow response (default), 1 = level-synchronous wavefronts, 2 / 3 = dataflow with
    whole warps / single-lane workers."""
    b = scenes.dense_lattice(20)           # 8000 bodies
    p = oracle.default_params(g=(0, 0, 0))
    want = b.copy()
    for _ in range(10):
        oracle.integrate(want, p)
        oracle.resolve_list(want, oracle.contacts(want))
    got, contacts, st = gpu_run(tr, b, 10, g=(0, 0, 0), response_mode=mode)
    check_state(got, want, "dense lattice")
    assert np.array_equal(contacts, oracle.contacts(want))
    assert st.last_response_rounds >= 10   # wavefronts (mode 1) / iterations of the longest dataflow worker (modes 0, 2, 3)


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_shuffled_ids_response_bit_exact(tr, oracle, mode):
    b = scenes.dense_lattice(16, shuffle_ids=True)
    p = oracle.default_params(g=(0, 0, 0))
    want = b.copy()
    for _ in range(6):
        oracle.integrate(want, p)
        oracle.resolve_list(want, oracle.contacts(want))
    got, _, _ = gpu_run(tr, b, 6, g=(0, 0, 0), response_mode=mode)
    check_state(got, want, "shuffled lattice")


def test_contact_capacity_overflow_is_loud(tr):
    b = scenes.dense_lattice(10)
    p = tr.default_params(max_contacts=16)
    with tr.World(p) as w:
        w.add(b)
        with pytest.raises(tr.TraceItError) as e:
            w.step(1)
        assert e.value.code == 5


# ---------------------------------------------------------------- pair gravity
@pytest.mark.parametrize("n", [1, 2, 300, 2048, 5000, 16384, 20001])
def test_gravity_force_vs_f64(tr, oracle, n):
    b = scenes.random_cloud(n, 1000 + n, spheres=True)
    p = tr.default_params(flags=tr.PAIR_GRAVITY, G=np.float32(1.0), g=(0, 0, 0))
    with tr.World(p) as w:
        w.add(b)
        f = w.debug_gravity().astype(np.float64)
    if n == 1:
        assert (f == 0).all()
        return
    f64 = oracle.pair_gravity_f64(b, 1.0)
    rel = np.linalg.norm(f - f64) / np.linalg.norm(f64)
    f32 = oracle.pair_gravity_f32(b, 1.0).astype(np.float64)
    rel_cpu = np.linalg.norm(f32 - f64) / np.linalg.norm(f64)
    print(f"n={n}: GPU rel L2 {rel:.3e}  (CPU fp32 restatement {rel_cpu:.3e})")
    assert rel <= FORCE_REL_L2_TOL


def test_gravity_coincident_and_zero_mass(tr, oracle):
    """Edge semantics of the restated formula: coincident DISTINCT bodies -> NaN (G*m*m/0 *
    normalize(0)), exactly like the CPU restatement; zero-mass bodies feel nothing but the
    self pair never contributes."""
    b = scenes.random_cloud(600, 3, spheres=True)
    b["pos"][1] = b["pos"][0]
    b["mass"][5] = 0
    p = tr.default_params(flags=tr.PAIR_GRAVITY, G=np.float32(1.0), g=(0, 0, 0))
    with tr.World(p) as w:
        w.add(b)
        f = w.debug_gravity()
    cpu = oracle.pair_gravity_f32(b, 1.0)
    assert np.array_equal(np.isnan(f), np.isnan(cpu))
    assert np.isnan(f[0]).all() and np.isnan(f[1]).all()
    assert (f[5] == 0).all()
    ok = ~np.isnan(cpu).any(axis=1)
    f64 = oracle.pair_gravity_f64(b, 1.0)
    assert np.linalg.norm(f[ok] - f64[ok]) / np.linalg.norm(f64[ok]) <= FORCE_REL_L2_TOL


def _traj_errors(got, want):
    extent = np.abs(want["pos"]).max()
    vrms = np.sqrt((want["vel"].astype(np.float64) ** 2).sum(axis=1).mean())
    epos = np.abs(got["pos"].astype(np.float64) - want["pos"]).max() / extent
    evel = np.abs(got["vel"].astype(np.float64) - want["vel"]).max() / vrms
    return epos, evel


@pytest.mark.parametrize("n,steps", [(8192, 100), (16384, 100)])
def test_gravity_trajectory_vs_restated_oracle(tr, oracle, n, steps):
    """Config 2 shape (gravity + drag + integration, collisions off), K = 100 steps, in the
    regime where fp32 trajectories are well conditioned (G = 1e-3: no hard two-body
    encounters within the horizon): max error vs the fp32 restated oracle <= 1e-4 of the
    cloud extent / rms speed."""
    b = scenes.random_cloud(n, scenes.SEED_CFG2, spheres=True)
    G = np.float32(1e-3)
    got, _, st = gpu_run(tr, b, steps, flags=tr.PAIR_GRAVITY, G=G, g=(0, 0, 0))
    want = b.copy()
    oracle.step(want, oracle.default_params(flags=oracle.PAIR_GRAVITY, G=G, g=(0, 0, 0)), steps)
    epos, evel = _traj_errors(got, want)
    print(f"n={n} K={steps} G=1e-3: max pos err / extent = {epos:.3e}, max vel err / vrms = {evel:.3e}")
    assert epos <= TRAJ_REL_TOL and evel <= TRAJ_REL_TOL
    assert st.interactions == n * n * steps


def _rms_errors(a, b):
    extent = np.abs(b["pos"]).max()
    vrms = np.sqrt((b["vel"].astype(np.float64) ** 2).sum(axis=1).mean())
    dp = a["pos"].astype(np.float64) - b["pos"]
    dv = a["vel"].astype(np.float64) - b["vel"]
    return np.sqrt((dp ** 2).sum(axis=1).mean()) / extent, np.sqrt((dv ** 2).sum(axis=1).mean()) / vrms


@pytest.mark.parametrize("n,steps", [(8192, 100), (65536, 4)])
def test_gravity_trajectory_cfg2_vs_fp64_yardstick(tr, oracle, n, steps):
    """Config 2 as the survey states it (G = 1, masses 1..10): hard close encounters make the
    K-step trajectory chaotic in fp32 - the CPU fp32 restatement itself drifts from the fp64
    restatement by up to several percent (max norm) at K = 100.  The honest bar there is the
    fp64 yardstick: the GPU must be at least as close to it as the CPU fp32
    restatement is (x1.5 slack), in rms position and velocity error."""
    b = scenes.random_cloud(n, scenes.SEED_CFG2, spheres=True)
    G = np.float32(1.0)
    got, _, _ = gpu_run(tr, b, steps, flags=tr.PAIR_GRAVITY, G=G, g=(0, 0, 0))
    w32 = b.copy()
    oracle.step(w32, oracle.default_params(flags=oracle.PAIR_GRAVITY, G=G, g=(0, 0, 0)), steps)
    w64 = b.copy()
    oracle.step(w64, oracle.default_params(flags=oracle.PAIR_GRAVITY | oracle.GRAVITY_F64, G=G, g=(0, 0, 0)), steps)
    gp, gv = _rms_errors(got, w64)
    cp, cv = _rms_errors(w32, w64)
    print(f"n={n} K={steps} G=1: rms err vs fp64  GPU pos {gp:.2e} vel {gv:.2e} | CPU fp32 pos {cp:.2e} vel {cv:.2e}")
    assert gp <= 1.5 * cp + 1e-7 and gv <= 1.5 * cv + 1e-7


def test_gravity_plus_collisions_small(tr, oracle):
    """Full step (gravity + drag + integrate + collisions) on a dense small cloud: the contact
    sets must agree exactly whenever positions do, and trajectories within tolerance."""
    b = scenes.random_cloud(3000, 17, spheres=True, density_l=0.6)
    flags = tr.PAIR_GRAVITY | tr.COLLISIONS
    got, contacts, _ = gpu_run(tr, b, 5, flags=flags, G=np.float32(1e-3), g=(0, 0, 0))
    want = b.copy()
    oracle.step(want, oracle.default_params(flags=oracle.PAIR_GRAVITY | oracle.COLLISIONS, G=np.float32(1e-3),
                                           g=(0, 0, 0)), 5)
    epos, evel = _traj_errors(got, want)
    assert epos <= TRAJ_REL_TOL and evel <= TRAJ_REL_TOL
    # the GPU's own contact set equals brute force on the GPU's own centres
    assert np.array_equal(contacts, oracle.contacts(got))


def test_sharded_world_of_one_rank_is_bit_identical(tr):
    b = scenes.random_cloud(9000, 4, spheres=True)
    p = tr.default_params(flags=tr.PAIR_GRAVITY, G=np.float32(1.0), g=(0, 0, 0))
    a, _, _ = gpu_run(tr, b, 3, flags=tr.PAIR_GRAVITY, G=np.float32(1.0), g=(0, 0, 0))
    with tr.World(p, rank=0, nranks=1) as w:
        w.add(b)
        assert w.owned_range() == (0, 9000)
        w.step(3)
        c = w.read_bodies()
    check_state(a, c, "sharded(1)")


def test_empty_world_and_errors(tr):
    with tr.World() as w:
        w.step(3)
        assert len(w) == 0
        assert len(w.contacts()) == 0
        with pytest.raises(tr.TraceItError):
            w.read_state(0, 5)
    with pytest.raises(tr.TraceItError):
        tr.World(tr.default_params(grid_bits=1))
    with pytest.raises(tr.TraceItError):
        tr.World(device=99)


# ---------------------------------------------------------------- full-size properties (BASELINE sizes)
def test_full_size_1m_properties(tr, oracle):
    """Config 3 at N = 1,048,576: (a) sampled targets' force vs fp64; (b) total momentum change
    ~ 0 (Newton's third law over the fp32 sums); (c) contact set equals the CPU grid
    restatement and, for sampled rows, brute force."""
    n = 1 << 20
    b = scenes.cfg3_cloud(n)
    p = tr.default_params(flags=tr.PAIR_GRAVITY | tr.COLLISIONS, G=np.float32(1.0), g=(0, 0, 0))
    with tr.World(p) as w:
        w.add(b)
        f = w.debug_gravity().astype(np.float64)
        idx = np.random.RandomState(0).choice(n, 128, replace=False).astype(np.int32)
        f64 = oracle.pair_gravity_f64_subset(b, 1.0, idx)
        rel = np.linalg.norm(f[idx] - f64) / np.linalg.norm(f64)
        print(f"1M sampled-force rel L2 vs fp64: {rel:.3e}")
        assert rel <= FORCE_REL_L2_TOL
        tot = np.abs(f.sum(axis=0)).max() / np.abs(f).sum(axis=0).max()
        print(f"1M |sum F| / sum |F| = {tot:.3e}")
        assert tot < 1e-5
        w.step(1)
        got = w.read_bodies()
        contacts = w.contacts()
        g = w.debug_grid()
    # contact set through the CPU grid on the GPU's own post-step centres
    _, cont, _, nk = oracle.grid_pairs(got, np.ones(n, np.uint8), (0, 0, 0), g["inv_cell"], g["bits"], want_candidates=False)
    assert np.array_equal(contacts, cont)
    # brute-force rows for a few sampled bodies (numpy fp32, same op order as the predicate)
    c = got["center"]
    for i in np.random.RandomState(1).choice(n, 8, replace=False):
        d = c[i] - c
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        rs = got["radius"][i] + got["radius"]
        hit = np.nonzero((d2 <= rs * rs))[0]
        hit = hit[hit != i]
        mine = set(map(tuple, contacts[(contacts[:, 0] == i) | (contacts[:, 1] == i)].tolist()))
        assert mine == {(min(i, j), max(i, j)) for j in hit.tolist()}


@pytest.mark.parametrize("crowd", [1500, 2600])
def test_crowded_cell(tr, oracle, crowd):
    """1500 / 2600 bodies in ONE cell (1.1M / 3.4M contacts): the counting sort, the on-demand stable order (a crowded
    cell is sorted by a CTA: in shared memory up to 2048 occupants, in global memory beyond) and the narrow phase
    must still produce the exact keys, order, candidates and contact set; every crowded body has thousands of
    contacts, so the K6 prep sorts its contact list too (2600: by the whole grid)."""
    n = 6000
    b = scenes.random_cloud(n, 5150, spheres=True, density_l=0.8)
    rs = np.random.RandomState(3)
    # the crowd is packed into a ball much smaller than one cell (cell edge ~1.008)
    b["pos"][:crowd] = (rs.uniform(-0.05, 0.05, size=(crowd, 3)) + np.array([0.5, 0.5, 0.5])).astype(np.float32)
    b["center"] = b["pos"]
    b["vel"][:crowd] = 0
    p = tr.default_params(g=(0, 0, 0), max_contacts=4_000_000)
    with tr.World(p) as w:
        w.add(b)
        g = w.debug_grid()
        ncand, cand = w.debug_candidates()
        w.step(1)
        got = w.read_bodies()
        contacts = w.contacts()
    keys = oracle.grid_keys(b, (0, 0, 0), g["inv_cell"], g["bits"])
    assert np.array_equal(g["keys"], keys)
    assert np.bincount(keys).max() >= crowd
    order, _ = oracle.grid_sort(keys, np.arange(n, dtype=np.int32), g["bits"])
    assert np.array_equal(g["sorted_ids"].astype(np.int32), order)
    ocand, _, nc, _ = oracle.grid_pairs(b, np.ones(n, np.uint8), (0, 0, 0), g["inv_cell"], g["bits"], want_contacts=False)
    assert ncand == nc and np.array_equal(cand, ocand)
    want = b.copy()
    op = oracle.default_params(g=(0, 0, 0))
    oracle.integrate(want, op)
    oc = oracle.contacts(want)
    oracle.resolve_list(want, oc)
    assert len(oc) > 1_000_000
    assert np.array_equal(contacts, oc)
    check_state(got, want, "crowded cell")


def test_full_size_dense_1m_bit_exact(tr, oracle):
    """Config 5 at N = 1,048,576 packed spheres (~3.1M contacts per step, ~800 response
    wavefronts): two full steps must equal the oracle (CPU grid contact set + ordered replay of
    the reference's impulse code) bit for bit."""
    n = 1 << 20
    b = scenes.cfg5_dense(n)
    want = b.copy()
    op = oracle.default_params(g=(0, 0, 0))
    with tr.World(tr.default_params(g=(0, 0, 0))) as w:
        w.add(b)
        for _ in range(2):
            w.step(1)
            got = w.read_bodies()
            contacts = w.contacts()
            g = w.debug_grid()
            oracle.integrate(want, op)
            _, oc, _, nk = oracle.grid_pairs(want, np.ones(n, np.uint8), (0, 0, 0), g["inv_cell"], g["bits"],
                                             want_candidates=False)
            oracle.resolve_list(want, oc)
            assert nk > 2_000_000
            assert np.array_equal(contacts, oc)
            check_state(got, want, "dense 1M")


def test_legacy_default_stream(tr, oracle):
    """tr_world_set_stream(NULL) = the legacy default stream (cannot be graph-captured)."""
    b = scenes.random_cloud(5000, 12, spheres=True, density_l=0.6)
    want = b.copy()
    oracle.step(want, oracle.default_params(), 2)
    with tr.World(tr.default_params()) as w:
        w.set_stream(None)
        w.add(b)
        w.step(2)
        got = w.read_bodies()
    check_state(got, want, "legacy stream")


def test_full_size_4m_sampled_forces(tr, oracle):
    """Config 4's body count on ONE GPU: sampled targets of a 4,194,304-body gravity pass vs fp64."""
    n = 1 << 22
    b = scenes.cfg4_cloud(n)
    p = tr.default_params(flags=tr.PAIR_GRAVITY, G=np.float32(1.0), g=(0, 0, 0))
    with tr.World(p) as w:
        w.add(b)
        f = w.debug_gravity().astype(np.float64)
    idx = np.random.RandomState(4).choice(n, 64, replace=False).astype(np.int32)
    f64 = oracle.pair_gravity_f64_subset(b, 1.0, idx)
    rel = np.linalg.norm(f[idx] - f64) / np.linalg.norm(f64)
    print(f"4M sampled-force rel L2 vs fp64: {rel:.3e}")
    assert rel <= FORCE_REL_L2_TOL
    assert np.isfinite(f).all()


def test_step_host_reclassifies_when_colliders_change(tr, oracle):
    """The per-frame host path skips the grid classification unless radii / shapes / live-ness
    changed; changing them between frames (as a UI edit of w.objects would) must take effect."""
    b = scenes.random_cloud(3000, 31, spheres=True, density_l=0.6)
    want = b.copy()
    cur = b.copy()
    op = oracle.default_params()
    with tr.World(tr.default_params()) as w:
        for frame in range(4):
            if frame == 2:
                for arr in (cur, want):
                    arr["radius"][5] = 30.0      # becomes an oversize sphere: side list
                    arr["isSphere"][9] = 0       # becomes a box
                    arr["size"][9] = (2.0, 0.3, 2.0)
                    arr["mass"][11] = 0.0        # drops out of every pair
            w.step_host(cur, 1)
            oracle.step(want, op, 1)
            check_state(cur, want, f"frame {frame}")
            assert np.array_equal(w.contacts(), oracle.contacts(want))
        assert w.stats().side_list_size >= 1   # the oversize sphere (+ bodies the reference's blow-up turned into NaN)


def test_bodies_added_between_steps_and_external_forces(tr, oracle):
    """World growth (push_back between frames), tr_set_force (obj.force = ...) and
    tr_write_bodies (direct edits of w.objects[i]) against the oracle doing the same."""
    a = scenes.random_cloud(1500, 41, spheres=True, density_l=0.6)
    extra = scenes.random_cloud(700, 42, spheres=True, density_l=0.5)
    op = oracle.default_params()
    with tr.World(tr.default_params()) as w:
        first = w.add(a)
        assert first == 0
        w.step(2)
        want = a.copy()
        oracle.step(want, op, 2)
        # grow the world past its initial capacity
        assert w.add(extra) == 1500
        want = np.concatenate([want, extra])
        w.set_force(3, (50.0, -20.0, 5.0))
        want["force"][3] = (50.0, -20.0, 5.0)
        w.step(2)
        oracle.step(want, op, 2)
        got = w.read_bodies()
        check_state(got, want, "after growth")
        assert np.array_equal(w.contacts(), oracle.contacts(want))
        # edit a few bodies in place
        edit = got[100:110].copy()
        edit["vel"] = (1.0, 2.0, 3.0)
        edit["radius"] = 0.9
        w.write_bodies(edit, first=100)
        want[100:110] = edit
        w.step(1)
        oracle.step(want, op, 1)
        check_state(w.read_bodies(), want, "after edit")
        assert len(w) == 2200


def test_degenerate_radii_and_infinite_positions(tr, oracle):
    b = scenes.random_cloud(2000, 43, spheres=True, density_l=0.6)
    b["radius"][::9] = 0.0
    b["radius"][4] = -0.5          # (ra+rb)^2 is still a valid square in the reference predicate
    b["radius"][5] = np.inf
    b["pos"][6] = (np.inf, 0, 0)
    b["pos"][7] = (-np.inf, np.inf, 1e30)
    b["center"] = b["pos"]
    want = b.copy()
    oracle.step(want, oracle.default_params(), 2)
    got, contacts, _ = gpu_run(tr, b, 2)
    check_state(got, want, "degenerate")
    assert np.array_equal(contacts, oracle.contacts(want))


def test_trajectory_ring_matches_reference_history(tr, oracle):
    """object::trajectory (src/main.cpp:793-808): the last maxTrajectoryPoints post-integration
    positions of every non-stationary body, kept on the device across steps, ring wrap included,
    surviving a capacity growth."""
    b = scenes.random_cloud(1200, 51, spheres=True, density_l=0.6)
    b["stationary"][3] = 1
    extra = scenes.random_cloud(600, 52, spheres=True, density_l=0.5)
    op = oracle.default_params()
    hist = {i: [] for i in (0, 3, 7, 1199, 1300)}
    want = b.copy()
    with tr.World(tr.default_params()) as w:
        w.enable_trajectory(5)
        w.add(b)
        assert len(w.trajectory(0)) == 0
        for s in range(12):
            if s == 6:
                w.add(extra)                       # grows past the initial capacity of 1024 -> 2048
                want = np.concatenate([want, extra])
            w.step(1)
            oracle.step(want, op, 1)
            for i in hist:
                if i < len(want) and not want["stationary"][i]:
                    hist[i].append(want["center"][i].copy())   # pos right after integration
        for i, pts in hist.items():
            got = w.trajectory(i)
            exp = np.array(pts[-5:], dtype=np.float32).reshape(-1, 3)
            assert got.shape == exp.shape, (i, got.shape, exp.shape)
            assert_bits_equal(got, exp, f"trajectory of body {i}")
        assert len(w.trajectory(3)) == 0 and len(w.trajectory(1300)) == 5
        w.enable_trajectory(0)
        with pytest.raises(tr.TraceItError):
            w.trajectory(0)


def test_gravity_survey_parameterisation_huge_masses(tr, oracle):
    """SURVEY.md 8d, config 2 alternative: masses uniform in [1,10]*1e9 with the reference's own
    G = 6.6743e-11f (G*m ~ 0.07..0.7).  Intermediate products reach ~1e20: no overflow, same
    accuracy as the G = 1 form."""
    n = 8192
    b = scenes.random_cloud(n, scenes.SEED_CFG2, spheres=True, mass_scale=1e9)
    p = tr.default_params(flags=tr.PAIR_GRAVITY, g=(0, 0, 0))      # G stays the reference constant
    with tr.World(p) as w:
        w.add(b)
        f = w.debug_gravity().astype(np.float64)
        w.step(3)
        got = w.read_bodies()
    assert np.isfinite(f).all() and np.isfinite(got["pos"]).all()
    f64 = oracle.pair_gravity_f64(b, float(np.float32(6.6743e-11)))
    rel = np.linalg.norm(f - f64) / np.linalg.norm(f64)
    print(f"huge masses: rel L2 vs fp64 {rel:.3e}")
    assert rel <= FORCE_REL_L2_TOL
    want = b.copy()
    oracle.step(want, oracle.default_params(flags=oracle.PAIR_GRAVITY, g=(0, 0, 0)), 3)
    epos, evel = _traj_errors(got, want)
    assert epos <= TRAJ_REL_TOL and evel <= TRAJ_REL_TOL


def test_dense_lattice_long_horizon_drift(tr, oracle):
    """SURVEY.md 8d config 5 drift check: a 32^3 = 32,768-sphere packed lattice, K = 50 steps, the
    contact set compared EVERY step and positions/velocities at the end.  Because detection and the
    ordered response are bit-exact the drift against the reference's step is exactly zero (the
    reference's inverted approach test makes many bodies blow up to inf/NaN over this horizon -
    they must do so identically)."""
    b = scenes.dense_lattice(32)
    n = len(b)
    op = oracle.default_params(g=(0, 0, 0))
    want = b.copy()
    ones = np.ones(n, np.uint8)
    with tr.World(tr.default_params(g=(0, 0, 0))) as w:
        w.add(b)
        g = w.debug_grid()
        for s in range(50):
            w.step(1)
            oracle.integrate(want, op)
            _, oc, _, _ = oracle.grid_pairs(want, ones * np.isfinite(want["center"]).all(axis=1), (0, 0, 0), g["inv_cell"], g["bits"],
                                            want_candidates=False)
            if not np.isfinite(want["center"]).all():
                oc = oracle.contacts(want)           # NaN/inf centres sit in the side list: brute force is the arbiter
            oracle.resolve_list(want, oc)
            assert np.array_equal(w.contacts(), oc), f"contact set differs at step {s}"
        got = w.read_bodies()
    check_state(got, want, "dense lattice K=50")
    finite = np.isfinite(want["pos"]).all(axis=1)
    drift = np.abs(got["pos"][finite].astype(np.float64) - want["pos"][finite]).max() if finite.any() else 0.0
    print(f"K=50 drift on {finite.sum()} finite bodies: {drift}")
    assert drift == 0.0


def test_parameter_changes_mid_run(tr, oracle):
    """tr_world_set_params between frames: grid geometry (re-allocation of the cell tables and
    re-capture of the broad-phase graph), collisions off/on, response executor, atmosphere and dt.
    The physics must follow the oracle driven with the same schedule."""
    b = scenes.random_cloud(6000, 61, spheres=True, density_l=0.6)
    want = b.copy()
    schedule = [
        dict(),                                            # defaults
        dict(grid_bits=3),                                 # coarser wrap: more aliasing, same contacts
        dict(grid_bits=8, grid_origin=(-3.5, 2.25, 0.75)), # bigger table, shifted origin
        dict(flags=0),                                     # collisions off
        dict(response_mode=1, atm_density=np.float32(0.4), dt=np.float32(0.01)),
        dict(grid_cell=np.float32(2.5)),                   # explicit (oversized) cell edge
    ]
    with tr.World(tr.default_params()) as w:
        w.add(b)
        for over in schedule:
            p = tr.default_params(**over)
            w.params = p
            w.step(2)
            op = oracle.default_params(flags=(oracle.COLLISIONS if p.flags & tr.COLLISIONS else 0),
                                       atm_density=p.atm_density, dt=p.dt)
            oracle.step(want, op, 2)
            got = w.read_bodies()
            check_state(got, want, f"after {over}")
            if p.flags & tr.COLLISIONS:
                assert np.array_equal(w.contacts(), oracle.contacts(want)), over


def test_all_box_world_uses_the_grid(tr, oracle):
    """The reference's creation form only ever builds AABB bodies (src/main.cpp:1302-1306): a world
    of 20,000 boxes of mixed sizes must go through the grid (empty side list), with the reference's
    AABB predicate and normal, bit for bit."""
    n = 20000
    b = scenes.random_cloud(n, 71, spheres=False, density_l=0.9)
    rs = np.random.RandomState(71)
    b["size"] = rs.uniform(0.2, 1.0, size=(n, 3)).astype(np.float32)
    b["isStatic"][::13] = 1
    want = b.copy()
    oracle.step(want, oracle.default_params(), 3)
    got, contacts, st = gpu_run(tr, b, 3)
    check_state(got, want, "box world")
    assert np.array_equal(contacts, oracle.contacts(want)) and len(contacts) > 1000
    assert st.side_list_size == 0


# ---------------------------------------------------------------- small worlds: the single-CTA step
@pytest.mark.parametrize("n,dens,steps", [(2, 0.3, 5), (37, 0.45, 40), (256, 0.5, 25), (256, 0.25, 6)])
def test_small_world_single_cta_step_bit_exact(tr, oracle, n, dens, steps):
    """Worlds of up to 256 bodies run whole steps in one kernel (integrate, brute-force detection with the
    reference's guards and predicates, sequential Gauss-Seidel replay): spheres, boxes, static, stationary and
    zero-mass bodies, pending forces - every field bit for bit, the contact list in the reference's order, and
    identical to the general grid pipeline on the same world."""
    rs = np.random.RandomState(n * 7 + steps)
    b = scenes.random_cloud(n, 100 + n, spheres=True, density_l=dens)
    b["isSphere"][::3] = 0
    b["size"][::3] = (0.6, 0.4, 0.5)
    b["isStatic"][::7] = 1
    b["stationary"][::5] = 1
    b["mass"][::13] = 0
    b["force"] = rs.uniform(-20, 20, size=(n, 3)).astype(np.float32)
    if n >= 37:      # degenerate bodies: the brute-force step must treat them exactly as the reference does
        b["pos"][5] = (np.nan, 0, 0)          # never touches anything (every comparison with NaN is false)
        b["pos"][6] = (1.0e30, -1.0e30, 0)    # far outside any grid domain
        b["radius"][8] = np.inf               # a sphere that touches every other sphere
        b["isSphere"][8] = 1
        b["center"] = b["pos"]
    want = b.copy()
    oracle.step(want, oracle.default_params(), steps)
    got, contacts, st = gpu_run(tr, b, steps)
    check_state(got, want, "single CTA")
    assert np.array_equal(contacts, oracle.contacts(want))
    assert st.kernel_launches <= 3, "a small world is one launch per tr_step call"
    got2, contacts2, st2 = gpu_run(tr, b, steps, debug_flags=tr.DEBUG_NO_TINY)
    check_state(got2, want, "general pipeline on a small world")
    assert np.array_equal(contacts2, contacts) and st2.kernel_launches > steps


def test_small_world_trajectory_ring_and_growth_past_the_limit(tr, oracle):
    """History ring on the single-CTA path, and a world that grows from 200 to 300 bodies mid-run (switching
    to the grid pipeline) against the oracle doing the same."""
    b = scenes.random_cloud(300, 5, spheres=True, density_l=0.5)
    want = b[:200].copy()
    with tr.World(tr.default_params()) as w:
        w.enable_trajectory(8)
        w.add(b[:200])
        w.step(12)
        oracle.step(want, oracle.default_params(), 12)
        hist = w.trajectory(3)
        assert hist.shape == (8, 3)
        w.add(b[200:])
        want = np.concatenate([want, b[200:]])
        w.step(6)
        oracle.step(want, oracle.default_params(), 6)
        check_state(w.read_bodies(), want, "grown world")
        assert np.array_equal(w.contacts(), oracle.contacts(want))
    # the ring holds the post-integration positions of the last 8 of the first 12 steps
    ref = b[:200].copy()
    pts = []
    for _ in range(12):
        tmp = ref.copy()
        oracle.step(tmp, oracle.default_params(flags=0), 1)      # integration only: the pushed point
        pts.append(tmp["pos"][3].copy())
        oracle.step(ref, oracle.default_params(), 1)
    assert_bits_equal(hist, np.array(pts[-8:]), "ring on the single-CTA path")


# ---------------------------------------------------------------- hub-shaped contact graphs
def test_hub_scene_slab_with_100k_boxes(tr, oracle):
    """Config 1's shape scaled up: the reference's 10000 x 5 x 10000 `earth` slab (main.cpp:1482-1491, stationary,
    not static) with 100,000 boxes inside its AABB at once.  Every box is in contact with the slab, so the slab's
    contact list has 10^5 entries (sorted by a CTA, not ranked entry against entry) and the reference's order makes
    its 10^5 contacts one sequential chain.  3 steps, bit for bit; the oracle replays the brute-force contact list."""
    n = 100_000
    rs = np.random.RandomState(9)
    b = scenes.random_cloud(n + 1, 91, spheres=False)
    b["pos"][1:, 0] = rs.uniform(-900, 900, n)
    b["pos"][1:, 1] = rs.uniform(-24.5, -15.5, n)
    b["pos"][1:, 2] = rs.uniform(-900, 900, n)
    b["size"][1:] = rs.uniform(0.3, 1.0, size=(n, 3)).astype(np.float32)
    b["mass"][1:] = rs.uniform(1, 10, n).astype(np.float32)
    b[0] = scenes.default_scene()[0]
    b["center"] = b["pos"]
    want = b.copy()
    op = oracle.default_params(flags=0)
    with tr.World(tr.default_params()) as w:
        w.add(b)
        for s in range(3):
            w.step(1)
            oracle.step(want, op, 1)                 # drag + uniform gravity + integration
            c = oracle.contacts(want)                # brute force, parallel over i
            oracle.resolve_list(want, c)             # == the sequential sweep
            assert np.array_equal(w.contacts(), c)
            if s == 0:   # (later the reference's inverted approach test has kicked most boxes out of the slab's AABB)
                assert (c[:, 0] == 0).sum() == n
        st = w.stats()
        check_state(w.read_bodies(), want, "hub scene")
    assert st.side_list_size >= 1   # the slab (+ boxes kicked out of the grid domain: far list)


def test_static_hub_orders_nothing(tr, oracle):
    """The same slab as an isStatic body: its contacts never change it, so they carry no dependency through it
    (the executor runs them concurrently) - and the result is still the sequential sweep's."""
    n = 30_000
    rs = np.random.RandomState(10)
    b = scenes.random_cloud(n + 1, 92, spheres=True, density_l=0.5)
    b[0] = scenes.default_scene()[0]
    b["isStatic"][0] = 1
    b["pos"][0] = (0, -30, 0)
    b["pos"][1:, 0] = rs.uniform(-60, 60, n)
    b["pos"][1:, 1] = rs.uniform(-34, -26, n)
    b["pos"][1:, 2] = rs.uniform(-60, 60, n)
    b["center"] = b["pos"]
    want = b.copy()
    oracle.step(want, oracle.default_params(), 4)
    got, contacts, st = gpu_run(tr, b, 4)
    check_state(got, want, "static hub")
    assert np.array_equal(contacts, oracle.contacts(want)) and (contacts[:, 0] == 0).sum() > n // 2
    assert st.last_response_rounds < 2000, "no chain through the static slab"


# ---------------------------------------------------------------- contact buffers grow on demand
def test_contact_buffers_grow_and_the_step_is_completed(tr, oracle):
    """32,768 packed spheres make ~100K contacts, more than the initial N + 32768 slots: the step that
    finds them skips its response on the device, the later steps of the same asynchronous batch become
    no-ops, and the next synchronisation grows the buffers and re-runs them - same bits as the oracle."""
    b = scenes.dense_lattice(32)
    want = b.copy()
    first = None
    for _ in range(3):
        oracle.step(want, oracle.default_params(flags=0, g=(0, 0, 0)), 1)
        c = oracle.contacts(want)
        oracle.resolve_list(want, c)
        first = len(c) if first is None else first
    assert first > 32768 + 32768, "the first step must not fit the initial buffers"
    with tr.World(tr.default_params(g=(0, 0, 0))) as w:
        w.add(b)
        for _ in range(3):
            w.step(1, sync=False)             # three steps in flight, the first one overflows
        got = w.read_bodies()
        assert np.array_equal(w.contacts(), c)
        assert w.stats().steps == 3
    check_state(got, want, "after overflow recovery")


@pytest.mark.parametrize("n_side,steps", [(64, 200), (101, 20)])
def test_response_modes_soak(tr, n_side, steps):
    """Soak of the version-stamped hand-off: a packed lattice (262,144 bodies x 200 steps; 1,030,301 x 20) stepped
    by the dataflow executors (single-lane and whole-warp workers) and by the level-synchronous executor, whose
    only synchronisation is a grid barrier per wavefront - the in-GPU arbiter where the CPU oracle is too slow.
    Every position and velocity must be bit-identical after every 10th step and at the end."""
    b = scenes.dense_lattice(n_side)
    worlds = [tr.World(tr.default_params(g=(0, 0, 0), response_mode=m)) for m in (1, 3, 2, 0)]
    try:
        for w in worlds:
            w.add(b)
        for s0 in range(0, steps, 10):
            for w in worlds:
                w.step(10, sync=False)
            ref_pos, ref_vel = worlds[0].read_state()
            for m, w in zip((3, 2, 0), worlds[1:]):
                pos, vel = w.read_state()
                assert_bits_equal(pos, ref_pos, f"step {s0 + 10}: mode {m} pos vs level-synchronous")
                assert_bits_equal(vel, ref_vel, f"step {s0 + 10}: mode {m} vel vs level-synchronous")
        assert worlds[0].stats().last_contacts > 0
    finally:
        for w in worlds:
            w.close()


def test_overflow_recovery_survives_a_long_asynchronous_batch(tr):
    """300 steps enqueued without a synchronisation, the FIRST of which overflows the initial contact buffers: the
    ring of per-step records (256 entries) has long wrapped when the host finally looks, the overflow notice has not.
    Same bits as a world that synchronised after every step (whose buffers grew at step 1, before anything else ran)."""
    b = scenes.dense_lattice(32)
    p = tr.default_params(g=(0, 0, 0))
    with tr.World(p) as w1, tr.World(p) as w2:
        w1.add(b)
        w2.add(b)
        for _ in range(300):
            w1.step(1, sync=False)
        for _ in range(300):
            w2.step(1)
        a, c = w1.read_bodies(), w2.read_bodies()
        for f in FIELDS:
            assert_bits_equal(a[f], c[f], f"long batch:{f}")
        assert np.array_equal(w1.contacts(), w2.contacts())
        assert w1.stats().steps == 300


def test_two_worlds_stepped_concurrently_from_two_host_threads(tr):
    """Two worlds on one GPU, each driven by its own host thread on its own stream: their collision steps overlap on
    the device (the K6 prep synchronises its grid by itself - it is a cooperative launch, so two of them cannot end up
    half resident waiting for each other).  Both must finish and equal a world stepped alone."""
    import threading
    b = scenes.dense_lattice(40)          # 64,000 packed spheres, ~190K contacts: the prep uses its whole grid
    p = tr.default_params(g=(0, 0, 0))
    steps = 40
    out, errs = {}, []

    def drive(name):
        try:
            with tr.World(p) as w:
                w.add(b)
                for _ in range(steps):
                    w.step(1, sync=False)
                out[name] = w.read_bodies()
        except Exception as e:      # surfaced below: an exception in a thread must fail the test
            errs.append(e)

    ts = [threading.Thread(target=drive, args=(k,)) for k in ("a", "b")]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in ts), "a world did not finish: the two collision steps deadlocked"
    assert not errs, errs
    with tr.World(p) as w:
        w.add(b)
        w.step(steps)
        alone = w.read_bodies()
    for name in ("a", "b"):
        for f in FIELDS:
            assert_bits_equal(out[name][f], alone[f], f"world {name}:{f}")


def test_narrow_phase_list_format_of_huge_worlds(tr, oracle):
    """Above 2^27 bodies a (candidate slot, owner lane) pair no longer fits one 32-bit list word and the narrow phase
    switches to a second instantiation (slot word + owner byte).  Forced on a small dense world here, incl. the
    candidate-listing debug mode and a mixed world: same contact sets, same state, bit for bit."""
    b = scenes.random_cloud(20000, 33, spheres=True, density_l=0.6)
    want = b.copy()
    oracle.step(want, oracle.default_params(), 2)
    got, contacts, _ = gpu_run(tr, b, 2, debug_flags=tr.DEBUG_BIG_LIST)
    check_state(got, want, "big list")
    assert np.array_equal(contacts, oracle.contacts(want)) and len(contacts) > 1000
    with tr.World(tr.default_params(debug_flags=tr.DEBUG_BIG_LIST)) as w, tr.World(tr.default_params()) as w0:
        w.add(b)
        w0.add(b)
        n1, c1 = w.debug_candidates()
        n0, c0 = w0.debug_candidates()
        assert n1 == n0 and np.array_equal(c1, c0)
    m = mixed_cloud()
    want = m.copy()
    oracle.step(want, oracle.default_params(), 3)
    got, contacts, _ = gpu_run(tr, m, 3, debug_flags=tr.DEBUG_BIG_LIST)
    check_state(got, want, "big list, mixed world")
    assert np.array_equal(contacts, oracle.contacts(want))


# ---------------------------------------------------------------- sparse steps: the one-cluster response kernel
def _slab_with_boxes(n, seed, static=False):
    rs = np.random.RandomState(seed)
    b = scenes.random_cloud(n + 1, 90 + seed, spheres=False)
    b["pos"][1:, 0] = rs.uniform(-900, 900, n)
    b["pos"][1:, 1] = rs.uniform(-24.5, -15.5, n)
    b["pos"][1:, 2] = rs.uniform(-900, 900, n)
    b["size"][1:] = rs.uniform(0.3, 1.0, size=(n, 3)).astype(np.float32)
    b["mass"][1:] = rs.uniform(1, 10, n).astype(np.float32)
    b[0] = scenes.default_scene()[0]
    b["isStatic"][0] = 1 if static else 0
    b["center"] = b["pos"]
    return b


@pytest.mark.parametrize("n,static", [(1500, True), (1500, False), (4000, True), (4090, True), (4096, True), (4100, True)])
def test_sparse_step_hub_and_the_limit(tr, oracle, n, static):
    """One body with ~n contacts.  A static slab orders nothing: the sparse-step kernel takes it up to its 4096-contact
    limit (straddled here: the boxes add a few contacts among themselves).  A movable slab is one chain of n contacts,
    more than the kernel ranks by linear scans: it leaves the step untouched - to the general kernels behind it in the
    graph (first step: nothing known yet), or, when the host was confident and left them out, to the host's re-run.
    Either way: the sequential sweep's bits, every step."""
    b = _slab_with_boxes(n, 3, static)
    want = b.copy()
    op = oracle.default_params(flags=0)
    with tr.World(tr.default_params()) as w, tr.World(tr.default_params(debug_flags=tr.DEBUG_NO_SPARSE)) as w2:
        w.add(b)
        w2.add(b)
        for s in range(4):
            w.step(1)
            w2.step(1)
            oracle.step(want, op, 1)
            c = oracle.contacts(want)
            oracle.resolve_list(want, c)
            assert np.array_equal(w.contacts(), c)
            if s == 0:
                assert (c[:, 0] == 0).sum() == n
                st = w.stats()   # nothing known before the first step: the sparse-step kernel tries, the general kernels stand behind it
                assert st.sparse_steps == (1 if static and len(c) <= 4096 else 0) and st.sparse_bailouts == 0
            check_state(w.read_bodies(), want, f"sparse hub step {s}")
            check_state(w2.read_bodies(), want, f"grid-wide hub step {s}")
        assert w2.stats().sparse_steps == 0 and w2.stats().sparse_bailouts == 0
        if not static:   # 1,500 contacts look sparse to the host, the slab's 1,500-entry list is not: one miss, then no more confidence
            assert w.stats().sparse_bailouts == 1


def _mixed(n, seed, density_l):
    b = scenes.random_cloud(n, seed, spheres=True, density_l=density_l)
    b["isSphere"][::3] = 0
    b["size"][::3] = (0.6, 0.4, 0.5)
    b["isStatic"][::7] = 1
    b["stationary"][::11] = 1
    b["mass"][::13] = 0
    b["rest"][::5] = 0.9
    return b


def test_sparse_and_dense_steps_alternate_in_one_run(tr, oracle):
    """A mixed cloud (spheres, boxes, static / stationary / massless bodies) that starts as a clump of 70,000 contacts and
    thins out to a few hundred.  Synchronous stepping: the first step is predicted sparse (nothing known yet), bails out
    and is re-run; the host then follows the records - general kernels while the clump lasts, the sparse-step kernel after.
    The same 200 steps as ONE asynchronous batch (nothing known for the whole batch: one bail-out, then general)."""
    b = _mixed(4000, 17, 0.25)
    want = b.copy()
    op = oracle.default_params()
    oracle.step(want, op, 1)
    assert len(oracle.contacts(want)) > 4096 * 4      # (detected after the response: still a clump)
    oracle.step(want, op, 199)
    assert len(oracle.contacts(want)) < 2048
    with tr.World(tr.default_params()) as w:
        w.add(b.copy())
        for _ in range(200):
            w.step(1)
        st = w.stats()
        check_state(w.read_bodies(), want, "sparse/dense run")
        assert st.sparse_steps >= 100 and st.steps == 200   # (step 1 overflows the default contact capacity: grown, re-run)
    with tr.World(tr.default_params()) as w:
        w.add(b.copy())
        w.step(200)
        st = w.stats()
        check_state(w.read_bodies(), want, "sparse/dense run, one batch")
        assert st.steps == 200


def test_sparse_prediction_fails_mid_batch(tr, oracle):
    """A thin cloud stepped a few times (the host now predicts sparse steps), then a clump of 20,000 contacts is dropped
    in and 30 steps are enqueued without synchronisation: the first of them bails out on the device, the 29 behind it are
    no-ops, and the host re-runs all 30 through the general kernels at the next synchronisation."""
    thin = scenes.random_cloud(3000, 5, spheres=True, density_l=0.8)
    clump = scenes.random_cloud(3000, 6, spheres=True, density_l=0.3)
    clump["pos"] += np.float32(500.0)
    clump["center"] = clump["pos"]
    want = thin.copy()
    op = oracle.default_params()
    with tr.World(tr.default_params()) as w:
        w.add(thin)
        for _ in range(5):
            w.step(1)
        oracle.step(want, op, 5)
        st = w.stats()
        assert st.sparse_steps == 5 and st.sparse_bailouts == 0
        w.add(clump)
        want = np.concatenate([want, clump])
        for _ in range(30):
            w.step(1, sync=False)
        oracle.step(want, op, 30)
        check_state(w.read_bodies(), want, "clump dropped in")
        st = w.stats()
        assert st.sparse_bailouts == 1 and st.steps == 35


@pytest.mark.parametrize("n,dens,steps", [(1 << 20, None, 300), (50_000, 2.0, 1000)])
def test_sparse_step_kernel_soak(tr, n, dens, steps):
    """Soak of the cluster kernel's version-stamped hand-off in distributed shared memory: the 1M cloud of config 3
    (~2,500 contacts per step) x 300 steps and a 50,000-body cloud x 1000 steps, each stepped by the sparse-step kernel
    and - the in-GPU arbiter where the CPU oracle is too slow - by the grid-wide ordering + dataflow executor.  Positions
    and velocities bit-identical after every 50th step, and the sparse kernel really ran."""
    b = scenes.cfg3_cloud(n) if dens is None else scenes.random_cloud(n, 23, spheres=True, density_l=dens)
    with tr.World(tr.default_params(g=(0, 0, 0))) as w, tr.World(tr.default_params(g=(0, 0, 0), debug_flags=tr.DEBUG_NO_SPARSE)) as w0:
        w.add(b)
        w0.add(b)
        for s0 in range(0, steps, 50):
            w.step(50, sync=False)
            w0.step(50, sync=False)
            pos, vel = w.read_state()
            pos0, vel0 = w0.read_state()
            assert_bits_equal(pos, pos0, f"step {s0 + 50}: pos, sparse-step kernel vs general kernels")
            assert_bits_equal(vel, vel0, f"step {s0 + 50}: vel, sparse-step kernel vs general kernels")
        st, st0 = w.stats(), w0.stats()
        assert st.steps == steps and st0.sparse_steps == 0
        assert st.sparse_steps >= steps * 3 // 4, (st.sparse_steps, st.sparse_bailouts)
        assert st.last_contacts == st0.last_contacts
