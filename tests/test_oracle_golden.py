"""Pin the CPU oracle (oracle/sparrow_oracle.c) against vectors produced by the live
reference's numba kernels (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, rel_err

SCENES = ["scene_cube05", "scene_c1", "scene_occluder", "scene_directional",
          "scene_canyon01", "scene_uneven", "scene_canyon015_dir"]


def test_rounding_model(oracle):
    g = load_golden("rounding_probes")
    # numba np.linalg.norm == x87 80-bit dnrm2 ; numba np.dot == FMA chain
    assert np.array_equal(oracle.vec_apply("sor_nrm3", g["v3"]), g["n3"])
    assert np.array_equal(oracle.vec_apply("sor_nrm2", g["v2"]), g["n2"])
    assert np.array_equal(oracle.vec_apply("sor_nrm3", g["l3"]), g["nl3"])
    assert np.array_equal(oracle.vec_apply("sor_dot3", g["a3"], g["b3"]), g["d3"])
    assert np.array_equal(oracle.vec_apply("sor_dot2", g["a2"], g["b2"]), g["d2"])
    # numpy (non-numba) norms used for the delay tables
    assert np.array_equal(oracle.vec_apply("sor_np_norm1d3", g["l3"]),
                          g["np_norm1d_l3"])
    assert np.array_equal(oracle.vec_apply("sor_np_norm_axis3", g["l3"]),
                          g["np_norm_axis1_l3"])


def test_tessellation(oracle):
    g = load_golden("tessellation")
    names = sorted(k[:-len("_walls")] for k in g if k.endswith("_walls"))
    assert len(names) == 5
    for n in names:
        pts, ids = oracle.process_patches(g[n + "_walls"], float(g[n + "_size"]))
        assert np.array_equal(pts, g[n + "_points"]), n
        assert np.array_equal(ids, g[n + "_ids"]), n
        assert np.array_equal(oracle.centers(pts), g[n + "_center"]), n
        assert np.array_equal(oracle.areas(pts), g[n + "_area"]), n


def test_visibility_predicates(oracle):
    g = load_golden("predicates")
    n = len(g["A"])
    vis = np.array([oracle.basic_visibility(g["A"][k], g["B"][k], g["S"][k], g["N"][k])
                    for k in range(n)])
    in_a = np.array([oracle.point_in_polygon(g["A"][k], g["S"][k], g["N"][k])
                     for k in range(n)])
    in_b = np.array([oracle.point_in_polygon(g["B"][k], g["S"][k], g["N"][k])
                     for k in range(n)])
    assert np.array_equal(in_a, g["inA"])
    assert np.array_equal(in_b, g["inB"])
    assert np.array_equal(vis, g["visible"])


def test_form_factor_pairs(oracle):
    g = load_golden("form_factor_pairs")
    ff = np.array([oracle.universal_form_factor(
        g["pts_i"][k], g["normal_i"][k], g["area_i"][k], g["pts_j"][k],
        g["normal_j"][k]) for k in range(len(g["ff"]))])
    nus = np.array([oracle.coincidence_check(g["pts_j"][k], g["pts_i"][k])
                    for k in range(len(g["ff"]))])
    assert np.array_equal(nus, g["nusselt"])
    assert np.max(np.abs(ff - g["ff"]) / np.abs(g["ff"])) < 1e-10


def test_point_patch(oracle):
    g = load_golden("point_patch")
    for mode in ("source", "receiver"):
        v = np.array([oracle.pt_solution(g["points"][k], g["patches"][k], mode)
                      for k in range(len(g["points"]))])
        ok = np.isfinite(g[mode])
        assert np.array_equal(np.isfinite(v), ok)
        assert np.max(np.abs(v[ok] - g[mode][ok])) <= 1e-13 * np.max(np.abs(g[mode][ok]))


@pytest.mark.parametrize("name", SCENES)
def test_scene_pipeline(oracle, name):
    g = load_golden(name)
    out = oracle.pipeline(
        g["walls_points"], g["walls_normal"], float(g["patch_size"]), g["source"],
        g["receivers"], float(g["speed_of_sound"]), float(g["dt"]),
        float(g["duration"]), int(g["max_order"]), g["air_attenuation"], g["vi"],
        g["vo"], g["brdf"].reshape(g["brdf"].shape[0], g["vi"].shape[1],
                                   g["vo"].shape[1], -1), g["brdf_index"],
        brdf_set_before_bake="brdf_dirs" in g)
    n = out["patches_center"].shape[0]
    # bit-exact: geometry, visibility, pair list, integer tables
    assert np.array_equal(out["patches_points"], g["patches_points"])
    assert np.array_equal(out["patches_center"], g["patches_center"])
    assert np.array_equal(out["patches_area"], g["patches_area"])
    vis_ref = np.unpackbits(g["visibility"])[:n * n].reshape(n, n).astype(bool)
    assert np.array_equal(out["visibility"], vis_ref)
    assert np.array_equal(out["visible_patches"], g["visible_patches"])
    assert out["visible_patches"].dtype == np.int32
    assert np.array_equal(out["pair_delays"][0::2], g["pair_delays"])
    assert np.array_equal(out["pair_delays"][1::2], g["pair_delays"])
    pairs = out["visible_patches"]
    p2o = g["p2o"].astype(np.int64)
    if "brdf_dirs" in g:       # directional BRDF set before bake -> p2o was baked
        assert np.array_equal(out["out_dir"][0::2], p2o[pairs[:, 0], pairs[:, 1]])
        assert np.array_equal(out["out_dir"][1::2], p2o[pairs[:, 1], pairs[:, 0]])
        both = vis_ref | vis_ref.T
        assert np.all(p2o[~both] == g["vo"].shape[1])
    else:
        assert np.all(p2o == 0) and np.all(out["out_dir"] == 0)
    assert np.array_equal(out["source_visibility"], g["source_visibility"])
    assert np.array_equal(out["receiver_visibility"], g["receiver_visibility"])
    assert np.array_equal(out["receiver_dir_index"], g["receiver_dir_index"])
    assert np.array_equal(out["receiver_delays"], g["receiver_delays"])
    d0 = out["distance_patches_to_source"]
    assert np.array_equal(d0, g["distance_patches_to_source"])
    assert np.array_equal(
        (d0 / float(g["speed_of_sound"]) / float(g["dt"])).astype(np.int64),
        g["source_delays"])
    # tolerance path
    assert rel_err(out["ff_pairs"], g["ff_pairs"]) < 1e-12
    assert int(g["ff_nnz_outside_pairs"]) == 0
    assert rel_err(out["energy_0"], g["energy_0"]) < 1e-12
    assert rel_err(out["energy_init_source"], g["energy_init_source"]) < 1e-12
    assert rel_err(out["receiver_factor"], g["receiver_factor"]) < 1e-12
    # dense tilde rows vs factored pair tables
    nd, nb = out["tilde_pairs"].shape[1:]
    dense = np.zeros((n, n, nd, nb))
    dense[pairs[:, 0], pairs[:, 1]] = out["tilde_pairs"][0::2]
    dense[pairs[:, 1], pairs[:, 0]] = out["tilde_pairs"][1::2]
    if "tilde" in g:
        ref_t = g["tilde"]
        if ref_t.shape[2:] == (1, 1) or ref_t.shape == dense.shape:
            assert rel_err(dense, ref_t.reshape(dense.shape)) < 1e-12
    else:
        assert rel_err(dense[g["tilde_rows"]], g["tilde_sample"]) < 1e-12
    etc = out["etc"]
    assert tuple(g["etc_shape"]) == etc.shape
    if "etc" in g:
        assert rel_err(etc, g["etc"]) < 1e-12
    else:
        assert rel_err(etc[g["etc_rows"]], g["etc_sample"]) < 1e-12
    assert rel_err(etc.sum(-1), g["etc_patch_sums"]) < 1e-12
    assert rel_err(out["etc_receiver_mono"], g["etc_receiver_mono"]) < 1e-12


def test_exchange_threads_bit_identical(oracle):
    """The time-sliced multi-thread oracle equals the serial one bit for bit."""
    g = load_golden("scene_occluder")
    kw = dict(air=g["air_attenuation"], vi=g["vi"], vo=g["vo"],
              brdf=g["brdf"].reshape(1, 1, 1, 1), brdf_index=g["brdf_index"],
              brdf_set_before_bake=False)
    args = (g["walls_points"], g["walls_normal"], float(g["patch_size"]), g["source"],
            g["receivers"], float(g["speed_of_sound"]), float(g["dt"]),
            float(g["duration"]), int(g["max_order"]))
    a = oracle.pipeline(*args, **kw, n_threads=1)["etc"]
    b = oracle.pipeline(*args, **kw, n_threads=4)["etc"]
    assert np.array_equal(a, b)
