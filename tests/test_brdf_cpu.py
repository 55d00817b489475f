"""BRDF builders (reference sparrowpy/brdf.py) -- restated reference tests
tests/test_brdf.py:13-140 without SOFA files: energy conservation, reciprocity,
diffuse limit, specular lobe placement, argument errors."""
import numpy as np
import numpy.testing as npt
import pytest

import sparrowpy_b200 as sp
from sparrowpy_b200 import brdf, pyfar_shim as pf


def hemisphere(n_az=24, n_col=12):
    """Equal-angle sampling of the upper hemisphere with solid-angle weights."""
    cols = (np.arange(n_col) + 0.5) * (np.pi / 2) / n_col
    az = np.arange(n_az) * 2 * np.pi / n_az
    c, a = np.meshgrid(cols, az, indexing="ij")
    xyz = np.stack([np.sin(c) * np.cos(a), np.sin(c) * np.sin(a), np.cos(c)], -1).reshape(-1, 3)
    w = (np.sin(c) * (np.pi / 2 / n_col) * (2 * np.pi / n_az)).reshape(-1)
    return pf.Coordinates.from_cartesian(xyz, weights=w)


def check_energy_conservation(receivers, data, absorption=0):
    w = np.asarray(receivers.weights).copy()
    w *= 2 * np.pi / np.sum(w)
    w *= np.cos(receivers.colatitude)
    energy = np.sum(np.real(data.freq) * w[..., np.newaxis], axis=1)
    npt.assert_almost_equal(energy, 1 - absorption, decimal=1)


def check_reciprocity(data):
    f = np.real(data.freq)
    npt.assert_almost_equal(f, np.transpose(f, (1, 0, 2)))


def test_diffuse_limit_is_one_over_pi():
    c = hemisphere()
    data = brdf.create_from_scattering(c, c, pf.FrequencyData(np.ones(3), [100, 200, 400]))
    assert data.freq.shape == (c.csize, c.csize, 3)
    npt.assert_almost_equal(np.real(data.freq), 1 / np.pi)
    check_energy_conservation(c, data)
    check_reciprocity(data)


@pytest.mark.parametrize("s", [0.0, 0.3])
def test_specular_lobe_and_energy(s):
    c = hemisphere()
    data = brdf.create_from_scattering(c, c, pf.FrequencyData(s + np.zeros(3), [100, 200, 400]))
    spec = c.copy()
    spec.azimuth = spec.azimuth + np.pi
    idx = c.find_nearest(spec)[0][0]
    npt.assert_array_less(0, np.real(data.freq)[np.arange(c.csize), idx])
    check_energy_conservation(c, data)
    check_reciprocity(data)


def test_absorption_scales_energy():
    c = hemisphere()
    data = brdf.create_from_scattering(
        c, c, pf.FrequencyData(0.3 + np.zeros(3), [100, 200, 400]),
        pf.FrequencyData(0.4 + np.zeros(3), [100, 200, 400]))
    check_energy_conservation(c, data, 0.4)


def test_directional_scattering():
    c = hemisphere(8, 4)
    w = np.asarray(c.weights) * 2 * np.pi / np.sum(c.weights)
    sd = np.tile((w * np.cos(c.colatitude) / np.pi)[None, :, None], (c.csize, 1, 2))
    data = brdf.create_from_directional_scattering(c, c, pf.FrequencyData(sd, [500, 1000]))
    npt.assert_almost_equal(np.real(data.freq), 1 / np.pi)       # Lambertian s_d


def test_argument_errors():
    c = hemisphere(4, 2)
    with pytest.raises(TypeError, match="scattering_coefficient"):
        brdf.create_from_scattering(c, c, np.ones(3))
    with pytest.raises(TypeError, match="source_directions"):
        brdf.create_from_scattering(np.zeros((3, 3)), c, pf.FrequencyData([1.0], [100]))
    with pytest.raises(TypeError, match="directional_scattering"):
        brdf.create_from_directional_scattering(c, c, pf.FrequencyData(np.ones((2, 2, 1)), [100]))
    with pytest.raises(NotImplementedError):
        brdf.create_from_scattering(c, c, pf.FrequencyData([1.0], [100]), file_path="x.sofa")


def test_builder_output_feeds_set_wall_brdf():
    rad = sp.DirectionalRadiosityFast.from_polygon(sp.testing.shoebox_room_stub(1, 1, 1), 1.0)
    c = hemisphere(4, 2)
    data = brdf.create_from_scattering(c, c, pf.FrequencyData([0.5, 0.6], [500, 1000]),
                                       pf.FrequencyData([0.1, 0.2], [500, 1000]))
    rad.set_wall_brdf(np.arange(6), data, c, c)
    assert rad.n_bins == 2 and rad._brdf[0].shape == (8, 8, 2)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_builders_match_the_live_reference(tag):
    """Vectors produced by the reference's own sparrowpy/brdf.py (driven with the pyfar
    stand-in, tests/golden/make_golden.py::gen_brdf_builders): both builders to 1e-14, and
    the reference's side effect on the receiver weights (normalised in place to 2 pi,
    brdf.py:103-104)."""
    from conftest import load_golden
    from sparrowpy_b200 import brdf, pyfar_shim as pf
    g = load_golden("brdf_builders")
    dirs, w, freqs = g[f"{tag}_dirs"], g[f"{tag}_weights"], g[f"{tag}_freqs"]
    s, alpha = g[f"{tag}_scattering"], g[f"{tag}_absorption"]
    src = pf.Coordinates.from_cartesian(dirs, weights=w.copy())
    rcv = pf.Coordinates.from_cartesian(dirs, weights=w.copy())
    b1 = brdf.create_from_scattering(src, rcv, pf.FrequencyData(s, freqs),
                                     pf.FrequencyData(alpha, freqs))
    np.testing.assert_allclose(np.real(b1.freq), g[f"{tag}_brdf_scattering"], rtol=1e-14)
    np.testing.assert_allclose(rcv.weights, g[f"{tag}_weights_after"], rtol=1e-15)
    rcv = pf.Coordinates.from_cartesian(dirs, weights=w.copy())
    b2 = brdf.create_from_scattering(src, rcv, pf.FrequencyData(s, freqs))
    np.testing.assert_allclose(np.real(b2.freq), g[f"{tag}_brdf_scattering_noabs"], rtol=1e-14)
    rcv = pf.Coordinates.from_cartesian(dirs, weights=w.copy())
    b3 = brdf.create_from_directional_scattering(
        src, rcv, pf.FrequencyData(g[f"{tag}_directional"], freqs),
        pf.FrequencyData(alpha, freqs))
    np.testing.assert_allclose(np.real(b3.freq), g[f"{tag}_brdf_directional"], rtol=1e-14)
    assert np.array_equal(b3.frequencies, freqs)
