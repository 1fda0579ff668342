"""The C++ side of the boundary on a B200: the headless harness (reference-shaped structs +
updatePhysics(world&, ofstream&) shim over the C ABI) must reproduce the verbatim reference's
1000-step default-scene run - state hash AND the trajectory dump file."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(ROOT, "traceit_b200", "host", "traceit_headless")


def test_headless_harness_default_scene(tmp_path):
    assert os.path.exists(HARNESS), "build it: make -C traceit_b200/host (or __graft_entry__.build())"
    r = subprocess.run([HARNESS, "--steps", "1000"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    g = load_golden("default_scene.npz")
    assert f"hash 0x{int(g['hash']):016x}" in r.stdout, r.stdout
    # the reference's side effect: "Координаты.txt" with the one-shot 300-point dumps
    got = (tmp_path / "Координаты.txt").read_text(encoding="utf-8")
    want = open(os.path.join(GOLDEN, "default_scene_coords.txt"), encoding="utf-8").read()
    assert got == want


def test_headless_harness_cloud_with_gravity(tmp_path):
    r = subprocess.run([HARNESS, "--steps", "5", "--cloud", "20000", "--gravity", "0.001", "--no-file"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bodies 20000 steps 5 hash 0x" in r.stdout
    assert not (tmp_path / "Координаты.txt").exists()


def test_shim_on_the_reference_own_types_equals_reference(tmp_path):
    """oracle/_ref/shim_vs_ref contains the UNMODIFIED reference TU and traceit_b200::updatePhysics
    instantiated on the reference's own world/object structs (INTEGRATION.md's three-line change).
    400 frames of the seeded world + projectiles + 300 sphere colliders through both: every body,
    every trajectory point, every transform and the Координаты.txt dumps must be identical."""
    exe = os.path.join(ROOT, "oracle", "_ref", "shim_vs_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/shim_vs_ref not built (needs /root/reference at build time)")
    r = subprocess.run([exe, "400", "300", str(tmp_path)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SHIM_EQUALS_REFERENCE" in r.stdout and "mismatching_bodies 0" in r.stdout
