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
