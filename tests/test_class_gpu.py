"""The drop-in class end to end on the GPU, driven exactly like the reference class
was when the golden fixtures were made (tests/golden/make_golden.py::run_scene)."""
import numpy as np
import pytest

from conftest import GPU_SCENES, load_golden, rel_err

pytestmark = pytest.mark.gpu
SCENES = GPU_SCENES
TOL = {"f64": 1e-6, "f32": 1e-4}


def run_class(g, dtype="f64"):
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf
    walls = [sp.Polygon(p, u, n) for p, u, n in
             zip(g["walls_points"], g["walls_up"], g["walls_normal"])]
    rad = sp.DirectionalRadiosityFast.from_polygon(walls, float(g["patch_size"]),
                                                   dtype=dtype)
    if "brdf_dirs" in g:
        coords = pf.Coordinates.from_cartesian(g["brdf_dirs"], weights=g["brdf_weights"])
        brdf, bidx = g["brdf"], g["brdf_index"]
        for m in range(brdf.shape[0]):
            rad.set_wall_brdf(np.nonzero(bidx == m)[0],
                              pf.FrequencyData(brdf[m] / np.pi, g["frequencies"]),
                              coords, coords)
        rad.set_air_attenuation(pf.FrequencyData(g["air_attenuation"], g["frequencies"]))
    rad.bake_geometry()
    rad.init_source_energy(pf.Coordinates(*g["source"]))
    rad.calculate_energy_exchange(float(g["speed_of_sound"]), float(g["dt"]),
                                  float(g["duration"]),
                                  max_reflection_order=int(g["max_order"]))
    return rad


@pytest.mark.parametrize("name", SCENES)
def test_pipeline_matches_reference(name):
    from sparrowpy_b200 import pyfar_shim as pf
    g = load_golden(name)
    rad = run_class(g)
    n = rad.n_patches
    vis_ref = np.unpackbits(g["visibility"])[:n * n].reshape(n, n).astype(bool)
    assert np.array_equal(rad.visibility_matrix, vis_ref)
    assert rad._visibility_matrix.dtype == bool
    assert np.array_equal(rad._visible_patches, g["visible_patches"])
    assert rad._visible_patches.dtype == np.int32
    p = g["visible_patches"]
    assert rel_err(rad.form_factors[p[:, 0], p[:, 1]], g["ff_pairs"]) < 1e-6
    assert np.count_nonzero(rad.form_factors) == np.count_nonzero(g["ff_pairs"])
    assert np.array_equal(rad._patch_2_brdf_outgoing_index, g["p2o"].astype(np.int64))
    assert np.array_equal(rad._source_visibility, g["source_visibility"])
    assert np.array_equal(rad._distance_patches_to_source, g["distance_patches_to_source"])
    assert rel_err(rad._energy_init_source, g["energy_init_source"]) < 1e-6
    tilde = rad._form_factors_tilde
    if "tilde" in g:
        assert tilde.shape == g["tilde"].shape
        assert rel_err(tilde, g["tilde"]) < 1e-6
    else:
        assert rel_err(tilde[g["tilde_rows"]], g["tilde_sample"]) < 1e-6
    etc = rad._energy_exchange_etc
    assert etc.shape == tuple(g["etc_shape"])
    if "etc" in g:
        assert rel_err(etc, g["etc"]) < 1e-6
    else:
        assert rel_err(etc[g["etc_rows"]], g["etc_sample"]) < 1e-6
    assert rel_err(etc.sum(-1), g["etc_patch_sums"]) < 1e-6
    rcv = pf.Coordinates.from_cartesian(g["receivers"])
    mono = rad.collect_energy_receiver_mono(rcv)
    assert mono.time.shape == g["etc_receiver_mono"].shape
    for r in range(mono.time.shape[0]):
        for b in range(mono.time.shape[1]):
            assert rel_err(mono.time[r, b], g["etc_receiver_mono"][r, b]) < 1e-6
    assert np.allclose(mono.times, np.arange(etc.shape[-1]) * float(g["dt"]))
    pw = rad.collect_energy_receiver_patchwise(rcv)
    assert pw.time.shape == (len(g["receivers"]), n, etc.shape[2], etc.shape[3])
    assert rel_err(pw.time.sum(-1), g["etc_receiver_patch_sums"]) < 1e-6
    # order 0 / recalculate semantics (RadiosityFast.py:547-555)
    rad.calculate_energy_exchange(float(g["speed_of_sound"]), float(g["dt"]),
                                  float(g["duration"]), max_reflection_order=0)
    assert rel_err(rad._energy_exchange_etc.sum(-1), g["etc_patch_sums"]) < 1e-6  # cached
    rad.calculate_energy_exchange(float(g["speed_of_sound"]), float(g["dt"]),
                                  float(g["duration"]), max_reflection_order=0,
                                  recalculate=True)
    assert rel_err(rad._energy_exchange_etc.sum(-1), g["etc_order0_sums"]) < 1e-6


def test_f32_histograms_within_1e4():
    from sparrowpy_b200 import pyfar_shim as pf
    g = load_golden("scene_directional")
    rad = run_class(g, dtype="f32")
    assert rel_err(rad._energy_exchange_etc[g["etc_rows"]], g["etc_sample"]) < 1e-4
    mono = rad.collect_energy_receiver_mono(pf.Coordinates.from_cartesian(g["receivers"]))
    for r in range(mono.time.shape[0]):
        for b in range(mono.time.shape[1]):
            assert rel_err(mono.time[r, b], g["etc_receiver_mono"][r, b]) < 1e-4


def test_reference_cube_visibility_pattern():
    """reference tests/test_DRadiosityFast.py:19-29"""
    import sparrowpy_b200 as sp
    rad = sp.DirectionalRadiosityFast.from_polygon(sp.testing.shoebox_room_stub(1, 1, 1), 0.5)
    rad.bake_geometry()
    np.testing.assert_array_equal(rad._visibility_matrix[:4, :4], False)
    np.testing.assert_array_equal(rad._visibility_matrix[:4, 4:], True)
    np.testing.assert_array_equal(rad._visibility_matrix[4:8, 4:8], False)
    np.testing.assert_array_equal(rad._visibility_matrix[4:8, 8:], True)
    # upper triangle only: 6 walls x 4 patches, same-wall pairs invisible
    assert np.sum(rad._visibility_matrix) == (24 * 24 - 6 * 16) // 2
    assert rad._visibility_matrix.shape == (24, 24)
    # row sums of the full (both triangles) form factor matrix ~ 1
    # (reference tests/test_universal_formfactor.py:140-157)
    ff = rad.form_factors
    area = rad.patches_area
    full = ff + (ff * area[:, None] / area[None, :]).T
    np.testing.assert_allclose(full.sum(axis=1), 1.0, atol=1e-2)


def test_diffuse_tilde_equals_form_factors():
    """reference tests/test_DRadiosityFast_order.py:20-46"""
    import sparrowpy_b200 as sp
    rad = sp.DirectionalRadiosityFast.from_polygon(sp.testing.shoebox_room_stub(2, 2, 2), 1.0)
    rad.bake_geometry()
    ff = rad.form_factors
    tilde = rad._form_factors_tilde[:, :, 0, 0]
    iu = np.triu_indices(rad.n_patches, 1)
    np.testing.assert_allclose(tilde[iu], ff[iu], rtol=1e-12)
    area = rad.patches_area
    np.testing.assert_allclose(tilde.T[iu], (ff * area[:, None] / area[None, :])[iu], rtol=1e-12)


def test_checkpoint_roundtrip_and_resume():
    """to_dict / from_dict (reference RadiosityFast.py:841-886, tests/
    test_DirectionalRadiosityFast.py:23-125): every array survives, and a resumed
    object collects the same receiver ETC without re-running anything."""
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf
    g = load_golden("scene_directional")
    rad = run_class(g)
    rcv = pf.Coordinates.from_cartesian(g["receivers"])
    mono = rad.collect_energy_receiver_mono(rcv).time
    d = rad.to_dict()
    assert set(d) >= {"walls_points", "visibility_matrix", "visible_patches", "form_factors",
                      "form_factors_tilde", "patch_2_brdf_outgoing_index",
                      "energy_exchange_etc", "energy_init_source", "speed_of_sound"}
    assert isinstance(d["form_factors"], list) and isinstance(d["n_patches"], int)
    rad2 = sp.DirectionalRadiosityFast.from_dict(d)
    for name in ("_visibility_matrix", "_visible_patches", "_form_factors",
                 "_form_factors_tilde", "_patch_2_brdf_outgoing_index",
                 "_energy_init_source", "_distance_patches_to_source",
                 "_energy_exchange_etc"):
        np.testing.assert_array_equal(np.asarray(getattr(rad2, name)),
                                      np.asarray(getattr(rad, name)), err_msg=name)
    assert rad2.speed_of_sound == rad.speed_of_sound
    mono2 = rad2.collect_energy_receiver_mono(rcv).time
    assert rel_err(mono2, mono) < 1e-12
    # an untouched object serialises its missing stages as the string 'None'
    fresh = sp.DirectionalRadiosityFast.from_polygon(sp.testing.shoebox_room_stub(1, 1, 1), 1.0)
    assert fresh.to_dict()["form_factors"] == "None"
    again = sp.DirectionalRadiosityFast.from_dict(fresh.to_dict())
    assert again.n_patches == fresh.n_patches and again._form_factors is None


def test_io_within_simulation_resumes_at_every_stage():
    """Mirror of the reference's test_io_within_simulation / test_io (tests/
    test_DirectionalRadiosityFast.py:23-125): a checkpoint taken after set-up, after the
    bake, after the source and after the exchange continues to the same ETC, and the
    round-tripped object compares equal (``__eq__``, RadiosityFast.py:875-879)."""
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf
    g = load_golden("scene_directional")
    c, dt, dur, order = float(g["speed_of_sound"]), float(g["dt"]), float(g["duration"]), 3
    src = pf.Coordinates(*g["source"])
    rcv = pf.Coordinates.from_cartesian(g["receivers"])
    walls = [sp.Polygon(p, u, n) for p, u, n in
             zip(g["walls_points"], g["walls_up"], g["walls_normal"])]
    rad = sp.DirectionalRadiosityFast.from_polygon(walls, float(g["patch_size"]))
    rt = sp.DirectionalRadiosityFast.from_dict(rad.to_dict())
    assert rt == rad and not (rt != rad)
    coords = pf.Coordinates.from_cartesian(g["brdf_dirs"], weights=g["brdf_weights"])
    for m in range(g["brdf"].shape[0]):
        rad.set_wall_brdf(np.nonzero(g["brdf_index"] == m)[0],
                          pf.FrequencyData(g["brdf"][m] / np.pi, g["frequencies"]),
                          coords, coords)
    rad.set_air_attenuation(pf.FrequencyData(g["air_attenuation"], g["frequencies"]))
    assert rt != rad
    stage0 = sp.DirectionalRadiosityFast.from_dict(rad.to_dict())      # before the bake
    assert stage0 == rad
    rad.bake_geometry()
    stage1 = sp.DirectionalRadiosityFast.from_dict(rad.to_dict())      # baked
    assert stage1 == rad and stage0 != rad
    rad.init_source_energy(src)
    stage2 = sp.DirectionalRadiosityFast.from_dict(rad.to_dict())      # baked + source
    assert stage2 == rad
    rad.calculate_energy_exchange(c, dt, dur, max_reflection_order=order)
    stage3 = sp.DirectionalRadiosityFast.from_dict(rad.to_dict())      # finished
    assert stage3 == rad and stage2 != rad
    assert rad != "something else"
    mono = rad.collect_energy_receiver_mono(rcv).time

    stage0.bake_geometry()
    stage0.init_source_energy(src)
    stage1.init_source_energy(src)
    for obj in (stage0, stage1, stage2, stage3):
        obj.calculate_energy_exchange(c, dt, dur, max_reflection_order=order, recalculate=True)
        assert rel_err(obj._energy_exchange_etc, rad._energy_exchange_etc) < 1e-13
        assert rel_err(obj.collect_energy_receiver_mono(rcv).time, mono) < 1e-12
    # a resumed object accepts further BRDF assignments (brdf_index is an array again)
    stage3.set_wall_brdf(np.array([0]), pf.FrequencyData(g["brdf"][0] / np.pi, g["frequencies"]),
                         coords, coords)
    assert int(stage3._brdf_index[0]) == len(stage3._brdf) - 1


def test_patchwise_collection_of_a_source_batch():
    """collect_energy_receiver_patchwise after init_source_energy_batch: (S, R, N, B, T),
    equal to the single-source results and summing to the mono ETC."""
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf
    g = load_golden("scene_directional")
    srcs = np.array([g["source"], [0.6, 1.5, 0.4]])
    rcv = pf.Coordinates.from_cartesian(g["receivers"])
    c, dt, dur, order = float(g["speed_of_sound"]), float(g["dt"]), float(g["duration"]), 2
    walls = [sp.Polygon(p, u, n) for p, u, n in
             zip(g["walls_points"], g["walls_up"], g["walls_normal"])]

    def fresh():
        r = sp.DirectionalRadiosityFast.from_polygon(walls, float(g["patch_size"]))
        coords = pf.Coordinates.from_cartesian(g["brdf_dirs"], weights=g["brdf_weights"])
        for m in range(g["brdf"].shape[0]):
            r.set_wall_brdf(np.nonzero(g["brdf_index"] == m)[0],
                            pf.FrequencyData(g["brdf"][m] / np.pi, g["frequencies"]),
                            coords, coords)
        r.set_air_attenuation(pf.FrequencyData(g["air_attenuation"], g["frequencies"]))
        r.bake_geometry()
        return r

    batch = fresh()
    batch.init_source_energy_batch(pf.Coordinates.from_cartesian(srcs))
    batch.calculate_energy_exchange(c, dt, dur, max_reflection_order=order)
    pw = batch.collect_energy_receiver_patchwise(rcv).time
    mono = batch.collect_energy_receiver_mono(rcv).time
    n_rcv = np.atleast_2d(g["receivers"]).shape[0]
    assert pw.shape == (2, n_rcv, batch.n_patches) + mono.shape[2:]
    assert rel_err(pw.sum(axis=2), mono) < 1e-12
    single = fresh()
    for s in range(2):
        single.init_source_energy(pf.Coordinates(*srcs[s]))
        single.calculate_energy_exchange(c, dt, dur, max_reflection_order=order,
                                         recalculate=True)
        assert rel_err(pw[s], single.collect_energy_receiver_patchwise(rcv).time) < 1e-13


def test_direct_sound_added_at_floor_delay():
    """collect_energy_receiver_mono(direct_sound=True) (RadiosityFast.py:594-657)."""
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf
    rad = sp.DirectionalRadiosityFast.from_polygon(sp.testing.shoebox_room_stub(3, 3, 3), 1.0)
    rad.bake_geometry()
    src, rcv = np.array([1.0, 1.0, 1.0]), np.array([[2.0, 2.2, 1.4]])
    rad.init_source_energy(pf.Coordinates(*src))
    rad.calculate_energy_exchange(343.2, 1e-3, 0.05, max_reflection_order=2)
    r = pf.Coordinates.from_cartesian(rcv)
    a = rad.collect_energy_receiver_mono(r, direct_sound=False).time
    b = rad.collect_energy_receiver_mono(r, direct_sound=True).time
    dist = np.linalg.norm(rcv[0] - src)
    k = int(dist / 343.2 / 1e-3)
    diff = b - a
    assert np.count_nonzero(diff) == 1
    np.testing.assert_allclose(diff[0, 0, k], 1 / (4 * np.pi * dist ** 2), rtol=1e-12)


def test_multi_source_batch_equals_single_sources():
    """init_source_energy_batch: S sources propagated in the same
    launches give exactly the S single-source results."""
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf
    g = load_golden("scene_directional")
    srcs = np.array([g["source"], [0.6, 1.5, 0.4], [2.5, 0.4, 1.7]])
    rcv = pf.Coordinates.from_cartesian(g["receivers"])
    c, dt, dur, order = float(g["speed_of_sound"]), float(g["dt"]), float(g["duration"]), 3

    def fresh():
        walls = [sp.Polygon(p, u, n) for p, u, n in
                 zip(g["walls_points"], g["walls_up"], g["walls_normal"])]
        rad = sp.DirectionalRadiosityFast.from_polygon(walls, float(g["patch_size"]))
        coords = pf.Coordinates.from_cartesian(g["brdf_dirs"], weights=g["brdf_weights"])
        for m in range(g["brdf"].shape[0]):
            rad.set_wall_brdf(np.nonzero(g["brdf_index"] == m)[0],
                              pf.FrequencyData(g["brdf"][m] / np.pi, g["frequencies"]),
                              coords, coords)
        rad.set_air_attenuation(pf.FrequencyData(g["air_attenuation"], g["frequencies"]))
        rad.bake_geometry()
        return rad

    rad = fresh()
    singles_etc, singles_mono = [], []
    for s in srcs:
        rad.init_source_energy(pf.Coordinates(*s))
        rad.calculate_energy_exchange(c, dt, dur, max_reflection_order=order, recalculate=True)
        singles_etc.append(rad._energy_exchange_etc.copy())
        singles_mono.append(rad.collect_energy_receiver_mono(rcv).time.copy())
    batch = fresh()
    batch.init_source_energy_batch(pf.Coordinates.from_cartesian(srcs))
    batch.calculate_energy_exchange(c, dt, dur, max_reflection_order=order)
    etc = batch._energy_exchange_etc
    assert etc.shape == (3,) + singles_etc[0].shape
    mono = batch.collect_energy_receiver_mono(rcv).time
    assert mono.shape == (3,) + singles_mono[0].shape
    for s in range(3):
        assert rel_err(etc[s], singles_etc[s]) < 1e-13
        assert rel_err(mono[s], singles_mono[s]) < 1e-12
    # order 0 (initial energy only) works for batches too
    batch.calculate_energy_exchange(c, dt, dur, max_reflection_order=0, recalculate=True)
    rad.init_source_energy(pf.Coordinates(*srcs[1]))
    rad.calculate_energy_exchange(c, dt, dur, max_reflection_order=0, recalculate=True)
    assert rel_err(batch._energy_exchange_etc[1], rad._energy_exchange_etc) < 1e-13


def test_ground_plane_multi_source_multi_receiver(oracle):
    """Config-3 analogue at test size (BASELINE.json: one ground plane, order 0,
    several sources and receivers; reference tests/
    test_DRadiosityFast_infinite_diffuse_plane.py:59-90): coplanar patches never see
    each other (P = 0), so the ETC is initial energy + receiver collection only."""
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf, scenes
    walls = scenes.ground_plane(-5, 5, -5, 5)
    rad = sp.DirectionalRadiosityFast.from_polygon([sp.Polygon(*w) for w in walls], 0.5)
    assert rad.n_patches == 400
    rad.bake_geometry()
    assert rad._visible_patches.shape == (0, 2) and not rad.visibility_matrix.any()
    rng = np.random.default_rng(0)
    srcs = np.column_stack([rng.uniform(-4, 4, 3), rng.uniform(-4, 4, 3), rng.uniform(1, 5, 3)])
    rcvs = np.column_stack([rng.uniform(-4, 4, 4), rng.uniform(-4, 4, 4), rng.uniform(1, 5, 4)])
    c, dt, dur = 343.2, 1e-3, 0.08
    rad.init_source_energy_batch(pf.Coordinates.from_cartesian(srcs))
    rad.calculate_energy_exchange(c, dt, dur, max_reflection_order=0)
    mono = rad.collect_energy_receiver_mono(pf.Coordinates.from_cartesian(rcvs)).time
    assert mono.shape == (3, 4, 1, 80)
    wp = np.array([w[0] for w in walls])
    wn = np.array([w[2] for w in walls])
    one = np.array([[[0.0, 0.0, 1.0]]])
    for s in range(3):
        ref = oracle.pipeline(wp, wn, 0.5, srcs[s], rcvs, c, dt, dur, 0, np.zeros(1), one, one,
                              np.full((1, 1, 1, 1), np.pi), np.zeros(1, np.int64))
        assert rel_err(mono[s], ref["etc_receiver_mono"]) < 1e-6
        assert rel_err(rad._energy_exchange_etc[s], ref["etc"]) < 1e-6


def test_source_directivity_scales_initial_energy_and_direct_sound():
    """A SoundSource with a directivity: one real factor per (patch, band) on the initial
    energy, the same for every outgoing direction (RadiosityFast.py:497-517), and one
    per (receiver, band) on the direct sound (:648-651)."""
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf, scenes, sound_object as so
    dirs4, w4 = scenes.hemisphere_directions(4, (45.0,))
    freqs = np.array([500.0, 2000.0])
    brdf = scenes.brdf_from_scattering(dirs4, w4, [0.5, 0.7], [0.1, 0.2])
    coords = pf.Coordinates.from_cartesian(dirs4, weights=w4)

    def make(source):
        rad = sp.DirectionalRadiosityFast.from_polygon(sp.testing.shoebox_room_stub(3, 2, 2),
                                                       0.5)
        rad.set_wall_brdf(np.arange(6), pf.FrequencyData(brdf, freqs), coords, coords)
        rad.set_air_attenuation(pf.FrequencyData([1e-3, 4e-3], freqs))
        rad.bake_geometry()
        rad.init_source_energy(source)
        return rad

    rng = np.random.default_rng(3)
    sphere = rng.normal(size=(60, 3))
    sphere /= np.linalg.norm(sphere, axis=1)[:, None]
    data = rng.uniform(0.2, 1.5, (60, 3))
    directivity = so.DirectivityMS.from_arrays(data, [400.0, 1000.0, 2500.0], sphere)
    pos = [1.1, 0.9, 1.2]
    omni = make(pf.Coordinates(*pos))
    src = so.SoundSource(pos, [1, 0, 0], [0, 0, 1], directivity=directivity)
    rad = make(src)
    e_omni, e_dir = omni._energy_init_source, rad._energy_init_source
    assert e_dir.shape == e_omni.shape == (rad.n_patches, 4, 2)
    fac = np.stack([src.get_directivity(rad.patches_center, f) for f in freqs], -1)
    assert fac.shape == (rad.n_patches, 2) and fac.min() > 0
    np.testing.assert_allclose(e_dir, e_omni * fac[:, None, :], rtol=1e-14)
    assert np.array_equal(rad._distance_patches_to_source, omni._distance_patches_to_source)
    # the exchange is linear in the initial energy: directivity scales what it feeds
    for r_ in (omni, rad):
        r_.calculate_energy_exchange(343.2, 0.5e-3, 0.02, max_reflection_order=2)
    rcv = pf.Coordinates.from_cartesian(np.array([[2.2, 1.3, 0.7], [0.4, 0.5, 1.6]]))
    d_omni, k_omni = omni.calculate_direct_sound(rcv)
    d_dir, k_dir = rad.calculate_direct_sound(rcv)
    fac_r = np.stack([src.get_directivity(rcv.cartesian, f) for f in freqs], -1)
    np.testing.assert_allclose(d_dir, d_omni * fac_r, rtol=1e-14)
    assert np.array_equal(k_dir, k_omni)
    etc = rad.collect_energy_receiver_mono(rcv, direct_sound=True).time
    assert etc.shape == (2, 2, 40) and np.isfinite(etc).all() and etc.sum() > 0
