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
    assert n_patches(scenes.city_block(), 0.5) == 100000             # config 5
    # bench.py's configurations build these scenes; sources / receivers stand in the street
    import bench
    for name, n in (("c1", 148), ("c2", 3700), ("c4", 19200), ("c5", 100000), ("c5s", 6280)):
        cfg = bench.CONFIGS[name]
        kind, arg = cfg["scene"]
        walls = (scenes.shoebox(*arg) if kind == "shoebox" else
                 scenes.city_block(0, arg) if kind == "city" else scenes.street_canyon(0, arg))
        assert n_patches(walls, cfg["patch"]) == n, name
        pts = np.array([w[0] for w in walls])
        lo, hi = pts.reshape(-1, 3).min(0), pts.reshape(-1, 3).max(0)
        for p in [cfg["source"]] + list(cfg["receivers"]):
            assert np.all(np.asarray(p) >= lo - 1e-9) and np.all(np.asarray(p) <= hi + 1e-9)
            if kind != "shoebox":                 # not inside a building (4 facades each)
                assert (len(walls) - 1) % 4 == 0
                for k in range(1, len(walls), 4):
                    q = np.concatenate([np.asarray(w[0]) for w in walls[k:k + 4]])
                    inside = (q[:, 0].min() < p[0] < q[:, 0].max() and
                              q[:, 1].min() < p[1] < q[:, 1].max())
                    assert not inside, (name, p)


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
