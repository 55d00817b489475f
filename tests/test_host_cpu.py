"""Host-side logic on CPU: tessellation vs the golden vectors, x87 emulation vs the
CPU's real long double, scene generators."""
import os
import subprocess

import numpy as np

from conftest import load_golden, REPO


def test_tessellation_matches_reference():
    from sparrowpy_b200 import geometry
    g = load_golden("tessellation")
    names = sorted(k[:-len("_walls")] for k in g if k.endswith("_walls"))
    for n in names:
        pts, ids = geometry.process_patches(g[n + "_walls"], float(g[n + "_size"]))
        assert np.array_equal(pts, g[n + "_points"]), n
        assert np.array_equal(ids, g[n + "_ids"]) and ids.dtype == np.int64
        assert np.array_equal(geometry.calculate_center(pts), g[n + "_center"]), n
        assert np.array_equal(geometry.calculate_area(pts), g[n + "_area"]), n


def test_x87_emulation_matches_long_double(tmp_path):
    exe = tmp_path / "x87_selftest"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-o", str(exe),
                           os.path.join(REPO, "tests", "native", "x87_selftest.cpp")])
    out = subprocess.run([str(exe), "100000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "0 mismatches" in out.stdout


def test_scene_sizes():
    from sparrowpy_b200 import geometry, scenes
    def n_patches(walls, size):
        return geometry.process_patches(np.array([w[0] for w in walls]), size)[0].shape[0]
    assert n_patches(scenes.shoebox(5, 6, 4), 1.0) == 148            # config 1
    assert n_patches(scenes.shoebox(5, 6, 4), 0.2) == 3700           # config 2
    assert n_patches(scenes.street_canyon(), 1.0) == 19200           # config 4
    assert n_patches(scenes.ground_plane(-50, 50, -50, 50), 0.5) == 40000   # config 3
    for w in scenes.street_canyon() + scenes.city_block():
        geometry.Polygon(*w)                                         # planarity / normal asserts


def test_visibility_predicate_shortcuts_match_oracle(tmp_path):
    """exact.cuh (the device predicate with its result-preserving shortcuts) compiled for
    the host and compared with the oracle's literal `_basic_visibility` on 800 k random,
    coplanar and lattice-degenerate cases."""
    obj = tmp_path / "sor.o"
    exe = tmp_path / "exact_selftest"
    flags = ["-O2", "-ffp-contract=off"]
    subprocess.check_call(["gcc", *flags, "-std=gnu11", "-c",
                           os.path.join(REPO, "oracle", "sparrow_oracle.c"), "-o", str(obj)])
    subprocess.check_call(["g++", *flags, "-o", str(exe),
                           os.path.join(REPO, "tests", "native", "exact_selftest.cpp"),
                           str(obj), "-lm"])
    out = subprocess.run([str(exe), "200000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert " 0 mismatches" in out.stdout
