"""Host-side behaviour of the drop-in class that needs no GPU: construction,
validation errors (reference RadiosityFast.py:211-337), BRDF bookkeeping."""
import numpy as np
import pytest

import sparrowpy_b200 as sp
from sparrowpy_b200 import pyfar_shim as pf


def test_init_from_polygon():
    """reference tests/test_DRadiosityFast.py:10-16"""
    rad = sp.DirectionalRadiosityFast.from_polygon(sp.testing.shoebox_room_stub(1, 1, 1), 0.2)
    assert rad.n_walls == 6 and rad.n_patches == 150
    assert rad.patches_points.shape == (150, 4, 3)
    assert rad.patches_area.shape == (150,)
    assert rad.patches_center.shape == (150, 3)
    assert rad.patches_size.shape == (150, 3)
    assert rad.patches_normal.shape == (150, 3)
    np.testing.assert_allclose(rad.walls_area, 1.0)


def test_check_raises_like_reference():
    walls = sp.testing.shoebox_room_stub(1, 1, 1)
    rad = sp.DirectionalRadiosityFast.from_polygon(walls, 0.5)
    kw = dict(walls_points=rad.walls_points, walls_normal=rad.walls_normal,
              walls_up_vector=rad.walls_up_vector, patches_points=rad.patches_points,
              n_patches=rad.n_patches, patch_to_wall_ids=rad._patch_to_wall_ids)
    sp.DirectionalRadiosityFast(**kw)
    with pytest.raises(ValueError, match="Normal of walls"):
        sp.DirectionalRadiosityFast(**{**kw, "walls_normal": rad.walls_normal[:3]})
    with pytest.raises(ValueError, match="patch_to_wall_ids"):
        sp.DirectionalRadiosityFast(**{**kw, "patch_to_wall_ids": np.zeros(24, int)})
    with pytest.raises(ValueError, match="form_factors need"):
        sp.DirectionalRadiosityFast(**{**kw, "form_factors": np.zeros((3, 3))})
    with pytest.raises(ValueError, match="Speed of sound"):
        sp.DirectionalRadiosityFast(**{**kw, "speed_of_sound": -1})
    with pytest.raises(ValueError, match="Air attenuation"):
        sp.DirectionalRadiosityFast(**{**kw, "air_attenuation": np.zeros(3),
                                       "frequencies": np.array([1.0])})


def test_set_wall_brdf_bookkeeping():
    """reference tests/test_DRadiosityFast.py:153-172: rotated directions point into
    the half space of the wall normal; brdf stored times pi."""
    rad = sp.DirectionalRadiosityFast.from_polygon(sp.testing.shoebox_room_stub(1, 1, 1), 1.0)
    dirs, w = sp.scenes.hemisphere_directions(4, (45.0,))
    coords = pf.Coordinates.from_cartesian(dirs, weights=w)
    brdf = sp.scenes.brdf_from_scattering(dirs, w, [0.5], [0.1])
    rad.set_wall_brdf(np.arange(6), pf.FrequencyData(brdf, [1000.0]), coords, coords)
    assert rad.n_bins == 1
    np.testing.assert_allclose(rad._brdf[0], brdf * np.pi)
    assert list(rad._brdf_index) == [0] * 6
    for i in range(6):
        out = rad._brdf_outgoing_directions[i].cartesian
        assert (out @ rad.walls_normal[i] > 0).all()
        np.testing.assert_allclose(np.linalg.norm(out, axis=-1), 1.0)
    with pytest.raises(AssertionError, match="Frequencies do not match"):
        rad.set_air_attenuation(pf.FrequencyData([0.0], [500.0]))
    with pytest.raises(AssertionError, match="positive half space"):
        bad = pf.Coordinates(0, 0, -1, weights=1)
        rad.set_wall_brdf([0], pf.FrequencyData(np.ones((1, 1, 1)), [1000.0]), bad, bad)


def test_argument_errors_before_any_gpu_work():
    rad = sp.DirectionalRadiosityFast.from_polygon(sp.testing.shoebox_room_stub(1, 1, 1), 1.0)
    with pytest.raises(ValueError, match="just one source"):
        rad.init_source_energy(pf.Coordinates([0.5, 0.2], [0.5, 0.2], [0.5, 0.2]))
    with pytest.raises(ValueError, match="direct_sound must be of type boolean"):
        rad.collect_energy_receiver_mono(pf.Coordinates(0.5, 0.5, 0.5), direct_sound=1)
    with pytest.raises(ValueError, match="must be of type pf.Coordinates"):
        rad.collect_energy_receiver_patchwise(np.zeros((1, 3)))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sparrowpy_b200._lib import SparrowB200Error
    rad = sp.DirectionalRadiosityFast.from_polygon(sp.testing.shoebox_room_stub(1, 1, 1), 1.0)
    with pytest.raises(SparrowB200Error, match="no CPU fallback"):
        rad.bake_geometry()


def test_equality_and_checkpoint_of_unbaked_objects():
    """``__eq__`` = equality of ``to_dict()`` (RadiosityFast.py:875-879) and the first two
    stages of the reference's test_io (tests/test_DirectionalRadiosityFast.py:23-48) --
    host-only state, so this runs without a GPU."""
    import numpy as np
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf
    walls = sp.testing.shoebox_room_stub(1, 1, 1)[:2]
    rad = sp.DirectionalRadiosityFast.from_polygon(walls, 1)
    out = sp.DirectionalRadiosityFast.from_dict(rad.to_dict())
    assert out == rad and not (out != rad)
    assert rad != 5 and rad != None  # noqa: E711
    freqs = [1000]
    rad.set_wall_brdf(np.arange(len(walls)), pf.FrequencyData(np.ones_like(freqs), freqs),
                      pf.Coordinates(0, 0, 1, weights=1), pf.Coordinates(0, 0, 1, weights=1))
    assert out != rad
    out = sp.DirectionalRadiosityFast.from_dict(rad.to_dict())
    assert out == rad
    assert isinstance(out._brdf_index, np.ndarray) and out._brdf_index.dtype == np.int64
    assert out._brdf_incoming_directions.dtype == object
    rad.set_air_attenuation(pf.FrequencyData(np.ones_like(freqs), freqs))
    assert out != rad
    out = sp.DirectionalRadiosityFast.from_dict(rad.to_dict())
    assert out == rad
    # a different patch size is a different object
    assert sp.DirectionalRadiosityFast.from_polygon(walls, 0.5) != rad
    # a resumed object accepts further BRDF assignments
    out.set_wall_brdf(np.array([1]), pf.FrequencyData(np.ones_like(freqs), freqs),
                      pf.Coordinates(0, 0, 1, weights=1), pf.Coordinates(0, 0, 1, weights=1))
    assert list(out._brdf_index) == [0, 1]


def test_direct_sound_matches_the_live_reference():
    """``calculate_direct_sound`` (RadiosityFast.py:605-657): vectors from the reference's own
    method (tests/golden/make_golden.py::gen_direct_sound) -- spreading loss and air
    attenuation to 1e-15, floor delay bins bit-exact.  Host arithmetic: runs without a GPU."""
    import numpy as np
    from conftest import load_golden
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf
    g = load_golden("direct_sound")
    rad = sp.DirectionalRadiosityFast.from_polygon(sp.testing.shoebox_room_stub(1, 1, 1), 1.0)
    freqs = np.array([250.0, 1e3, 4e3])
    rad.set_air_attenuation(pf.FrequencyData(g["air_attenuation"], freqs))
    rad._source = pf.Coordinates(*g["source"])
    rad._speed_of_sound = float(g["speed_of_sound"])
    rad._etc_time_resolution = float(g["dt"])
    direct, delay = rad.calculate_direct_sound(pf.Coordinates.from_cartesian(g["receivers"]))
    np.testing.assert_allclose(direct, g["direct_sound"], rtol=1e-15)
    assert np.array_equal(delay, g["n_sample_delay"])
    with pytest.raises(ValueError, match="pf.Coordinates"):
        rad.calculate_direct_sound(g["receivers"])


def test_wall_rotation_invariants_of_the_reference_tests():
    """``_rotate_coords_to_normal`` (RadiosityFast.py:971-986; pyfar's Orientations calls
    restated on scipy): the invariants the reference's tests state
    (tests/test_DRadiosityFast.py:153-172: all rotated directions lie in the wall's positive
    half space; :46-107: two orthogonal walls with flipped up vectors see each other under
    mirrored directions), the closed-form rotation, and the committed vectors."""
    import numpy as np
    from conftest import load_golden
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf
    from sparrowpy_b200.radiosity import _rotate_coords_to_normal
    g = load_golden("wall_rotation")
    coords = pf.Coordinates.from_cartesian(g["dirs"], weights=g["weights"])
    for n, u, want in zip(g["normals"], g["ups"], g["rotated"]):
        src, rcv = _rotate_coords_to_normal(n, u, coords, coords)
        assert src.cshape == coords.cshape and np.array_equal(src.weights, coords.weights)
        np.testing.assert_allclose(src.cartesian, want, atol=1e-15)
        np.testing.assert_allclose(rcv.cartesian, want, atol=1e-15)
        assert (src.cartesian @ n > 0).all()                       # positive half space
        np.testing.assert_allclose(src.radius, 1.0, atol=1e-15)
        m = pf.rotation_to_wall_frame(n, u)                        # normal <- z, up <- x
        np.testing.assert_allclose(src.cartesian, g["dirs"] @ m.T, atol=1e-14)
        np.testing.assert_allclose(m @ [0, 0, 1], n, atol=1e-15)
        np.testing.assert_allclose(m @ [1, 0, 0], u, atol=1e-15)
    # reference test_patch_2_out_dir_mapping: walls x = 0 (up +z) and y = 0 (up -z)
    samples = np.array([[np.cos(a) * np.sin(c), np.sin(a) * np.sin(c), np.cos(c)]
                        for c in np.deg2rad([0.0, 45.0, 90.0])
                        for a in np.deg2rad(np.arange(0, 360, 45))])
    samples[np.abs(samples) < 1e-15] = 0
    dirs = pf.Coordinates.from_cartesian(samples, weights=np.ones(len(samples)))
    o0, _ = _rotate_coords_to_normal([1, 0, 0], [0, 0, 1], dirs, dirs)
    o1, _ = _rotate_coords_to_normal([0, 1, 0], [0, 0, -1], dirs, dirs)
    c0, c1 = np.array([0, .5, .5]), np.array([.5, 0, .5])
    i0 = int(np.argmin(np.linalg.norm(o0.cartesian - (c1 - c0) / np.linalg.norm(c1 - c0), axis=1)))
    i1 = int(np.argmin(np.linalg.norm(o1.cartesian - (c0 - c1) / np.linalg.norm(c1 - c0), axis=1)))
    assert i0 == i1                         # symmetry over the x = y plane, up vectors flipped
    v0, v1 = o0.cartesian[i0], o1.cartesian[i1]
    np.testing.assert_allclose(v0[2], 0, atol=1e-7)
    np.testing.assert_allclose(v0, -v1, atol=1e-7)
    # non-perpendicular view / up is rejected like pf.Orientations.from_view_up
    with pytest.raises(ValueError, match="perpendicular"):
        _rotate_coords_to_normal([1, 0, 0], [1, 0, 1], dirs, dirs)
