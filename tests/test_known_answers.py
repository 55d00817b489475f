"""Known-answer tests in the style of the reference's own suite, run twice: against the
CPU oracle (always) and against the CUDA kernels through the C ABI (``-m gpu``).

They mirror reference tests/test_universal_formfactor.py (analytic form factors of
parallel and perpendicular rectangles, :12-135, tolerances 1.5 % / 2 % / 5 %; energy
conservation of a closed room, :137-157, 1e-2; reciprocity, :160-206) and
tests/test_visibility.py (:46-73 basic visibility, :75-207 matrix assembly of an
occluded scene).  The analytic expressions are the textbook configuration factors of
Howell's catalogue (C-11, C-14, C-15), written out here from the catalogue.
"""
import numpy as np
import pytest

from sparrowpy_b200 import geometry, scenes


# ---------------------------------------------------------------------------
# analytic configuration factors
# ---------------------------------------------------------------------------
def ff_parallel(a, b, c):
    """Identical, directly opposed rectangles a x b at distance c (catalogue C-11)."""
    x, y = a / c, b / c
    return 2.0 / (np.pi * x * y) * (
        0.5 * np.log((1 + x * x) * (1 + y * y) / (1 + x * x + y * y))
        + x * np.sqrt(1 + y * y) * np.arctan(x / np.sqrt(1 + y * y))
        + y * np.sqrt(1 + x * x) * np.arctan(y / np.sqrt(1 + x * x))
        - x * np.arctan(x) - y * np.arctan(y))


def ff_perpendicular_edge(w, h, edge):
    """Rectangle 1 (w x edge) to rectangle 2 (h x edge), perpendicular, sharing the edge
    (catalogue C-14)."""
    hh, ww = h / edge, w / edge
    s = hh * hh + ww * ww
    log_term = ((1 + ww * ww) * (1 + hh * hh) / (1 + s)
                * (ww * ww * (1 + s) / ((1 + ww * ww) * s)) ** (ww * ww)
                * (hh * hh * (1 + s) / ((1 + hh * hh) * s)) ** (hh * hh))
    return (ww * np.arctan(1 / ww) + hh * np.arctan(1 / hh)
            - np.sqrt(s) * np.arctan(1 / np.sqrt(s)) + 0.25 * np.log(log_term)) / (np.pi * ww)


def rect(origin, u, v):
    o, u, v = (np.asarray(x, float) for x in (origin, u, v))
    return np.array([o, o + u, o + u + v, o + v])


def normal_of(p):
    n = np.cross(p[1] - p[0], p[3] - p[0])
    return n / np.linalg.norm(n)


# ---------------------------------------------------------------------------
# implementations under test
# ---------------------------------------------------------------------------
class OracleImpl:
    name = "oracle"

    def __init__(self, oracle):
        self.o = oracle

    def ff(self, points, normals, pairs):
        areas = geometry.calculate_area(points)
        return self.o.ff_pairs(points, normals, areas, np.asarray(pairs, np.int64))

    def vis(self, centers, normals, points):
        return self.o.visibility_p2p(centers, normals, points).astype(bool)

    def vis_points(self, pts, centers, normals, points):
        return np.array([self.o.visibility_pt2p(p, centers, normals, points) for p in pts],
                        bool)


class CudaImpl:
    name = "cuda"

    @staticmethod
    def _t(a, dtype=None):
        import torch
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype or torch.float64).cuda()

    def ff(self, points, normals, pairs):
        import torch
        from sparrowpy_b200 import bake
        areas = geometry.calculate_area(points)
        ff, _ = bake.form_factors(self._t(points), self._t(normals), self._t(areas),
                                  self._t(np.asarray(pairs), torch.int32))
        return ff.cpu().numpy()

    def vis(self, centers, normals, points):
        from sparrowpy_b200 import bake
        return bake.visibility_p2p(self._t(centers), self._t(normals),
                                   self._t(points)).cpu().numpy()

    def vis_points(self, pts, centers, normals, points):
        from sparrowpy_b200 import bake
        return bake.visibility_pt2p(self._t(np.asarray(pts, float)), self._t(centers),
                                    self._t(normals), self._t(points)).cpu().numpy()


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def impl(request, oracle):
    return OracleImpl(oracle) if request.param == "oracle" else CudaImpl()


# ---------------------------------------------------------------------------
# form factors
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("height", [1.0, 2.0, 3.0, 4.0])
@pytest.mark.parametrize("distance", [1.0, 2.0, 3.0, 4.0])
def test_parallel_facing_patches(impl, height, distance):
    """reference tests/test_universal_formfactor.py:12-46 (1.5 %)"""
    width = 1.0
    p1 = rect([0, 0, 0], [width, 0, 0], [0, 0, height])              # normal -y ... +y below
    p1 = p1[[0, 1, 2, 3]]
    p2 = rect([0, distance, 0], [0, 0, height], [width, 0, 0])
    pts = np.array([p1, p2])
    nrm = np.array([[0.0, 1.0, 0.0], [0.0, -1.0, 0.0]])
    got = impl.ff(pts, nrm, [[0, 1]])[0]
    exact = ff_parallel(width, height, distance)
    assert abs(got - exact) / exact < 0.015


@pytest.mark.parametrize("width", [1.0, 2.0, 3.0])
@pytest.mark.parametrize("height", [1.0, 2.0, 3.0])
def test_perpendicular_patches_sharing_an_edge(impl, width, height):
    """reference tests/test_universal_formfactor.py:49-88 (2 %); the shared vertices make
    this the Nusselt-analogue branch (geometry.py:719-748)."""
    edge = 1.0
    p1 = rect([0, 0, 0], [0, edge, 0], [0, 0, height])               # in x = 0, normal +x
    p2 = rect([0, 0, 0], [width, 0, 0], [0, edge, 0])                # in z = 0, normal +z
    pts = np.array([p2, p1])
    nrm = np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0]])
    got = impl.ff(pts, nrm, [[0, 1]])[0]
    exact = ff_perpendicular_edge(width, height, edge)
    assert abs(got - exact) / exact < 0.02


def _room(x, y, z, patch):
    walls = scenes.shoebox(x, y, z)
    pts, ids = geometry.process_patches(np.array([w[0] for w in walls]), patch)
    nrm = np.array([w[2] for w in walls], float)[ids]
    return pts, nrm, ids


@pytest.mark.parametrize("dims", [(2.0, 1.0, 2.0), (3.0, 2.0, 2.0)])
def test_form_factor_energy_conservation(impl, dims):
    """closed room: sum_j F_ij = 1 within 1e-2 (reference
    tests/test_universal_formfactor.py:137-157)"""
    pts, nrm, ids = _room(*dims, 1.0)
    n = len(ids)
    pairs = np.array([(i, j) for i in range(n) for j in range(i + 1, n) if ids[i] != ids[j]])
    ff = impl.ff(pts, nrm, pairs)
    area = geometry.calculate_area(pts)
    rows = np.zeros(n)
    np.add.at(rows, pairs[:, 0], ff)
    np.add.at(rows, pairs[:, 1], ff * area[pairs[:, 0]] / area[pairs[:, 1]])   # reciprocity
    assert np.all(np.abs(1 - rows) < 1e-2)
    assert abs(n - rows.sum()) / n < 1e-2


@pytest.mark.parametrize("width", [1.0, 2.0, 4.0])
def test_reciprocity_of_unequal_patches(impl, width):
    """A_i F_ij = A_j F_ji: evaluating the pair in either order gives the reciprocal
    factors (reference tests/test_universal_formfactor.py:160-206, 1e-9 there on the
    dense tilde; the integration itself is symmetric to ~1e-6)."""
    p1 = rect([0, 0, 0], [1, 0, 0], [0, 0, 1])
    p2 = rect([0, 2, 0], [0, 0, 1], [width, 0, 0])
    pts = np.array([p1, p2])
    nrm = np.array([[0.0, 1.0, 0.0], [0.0, -1.0, 0.0]])
    f12 = impl.ff(pts, nrm, [[0, 1]])[0]
    f21 = impl.ff(pts[::-1].copy(), nrm[::-1].copy(), [[0, 1]])[0]
    assert abs(f12 * 1.0 - f21 * width) / (f12 * 1.0) < 1e-5


# ---------------------------------------------------------------------------
# visibility
# ---------------------------------------------------------------------------
def test_visibility_matrix_of_a_convex_room(impl):
    """every pair of patches on different walls sees each other, coplanar patches never
    do (reference geometry.py:903-906, tests/test_DRadiosityFast.py cube pattern)"""
    pts, nrm, ids = _room(2.0, 2.0, 2.0, 1.0)
    cen = geometry.calculate_center(pts)
    vis = impl.vis(cen, nrm, pts)
    n = len(ids)
    want = np.triu(ids[:, None] != ids[None, :], 1)
    assert vis.shape == (n, n) and np.array_equal(vis, want)


def test_visibility_with_an_occluding_plate(impl):
    """two facing walls with a plate in between: exactly the pairs whose connecting
    segment passes through the plate are blocked (reference tests/test_visibility.py:75-207
    assembles such matrices by hand)"""
    left = rect([0, 0, 0], [0, 4, 0], [0, 0, 4])        # x = 0, normal +x
    right = rect([4, 0, 0], [0, 0, 4], [0, 4, 0])       # x = 4, normal -x
    plate = rect([2, 1, 1], [0, 2, 0], [0, 0, 2])       # x = 2, 2 x 2 in the middle
    walls = np.array([left, right, plate])
    pts, ids = geometry.process_patches(walls, 1.0)
    nrm = np.array([normal_of(w) for w in walls])[ids]
    cen = geometry.calculate_center(pts)
    vis = impl.vis(cen, nrm, pts)
    n = len(ids)
    for i in range(n):
        for j in range(i + 1, n):
            if ids[i] == ids[j]:
                want = False
            elif {ids[i], ids[j]} == {0, 1}:
                # crossing point of the segment with the plane x = 2
                mid = 0.5 * (cen[i] + cen[j])
                if min(abs(mid[1] - 1), abs(mid[1] - 3), abs(mid[2] - 1), abs(mid[2] - 3)) < 1e-3:
                    continue    # on the plate's rim: pinned by the golden predicate vectors
                want = not (1 < mid[1] < 3 and 1 < mid[2] < 3)
            else:
                continue        # wall <-> plate pairs: checked against the oracle elsewhere
            assert vis[i, j] == want, (i, j)


def test_point_visibility_behind_a_plate(impl):
    """sources / receivers use the walls as blockers (RadiosityFast.py:482-486): a point
    in front of the plate sees the near wall's patches, the far wall's patches only where
    the plate does not cover them (reference tests/test_visibility.py:209-247)"""
    left = rect([0, 0, 0], [0, 4, 0], [0, 0, 4])
    right = rect([4, 0, 0], [0, 0, 4], [0, 4, 0])
    plate = rect([2, 1, 1], [0, 2, 0], [0, 0, 2])
    walls = np.array([left, right, plate])
    pts, ids = geometry.process_patches(walls, 1.0)
    wall_n = np.array([normal_of(w) for w in walls])
    cen = geometry.calculate_center(pts)
    src = np.array([[1.0, 2.0, 2.0]])
    vis = impl.vis_points(src, cen, wall_n, walls)[0]
    for k in range(len(ids)):
        if ids[k] == 0:
            assert vis[k]
        elif ids[k] == 1:
            # the ray to the far wall crosses x = 2 at src + (cen - src) / 3
            hit = src[0] + (cen[k] - src[0]) / 3.0
            if min(abs(hit[1] - 1), abs(hit[1] - 3), abs(hit[2] - 1), abs(hit[2] - 3)) < 1e-3:
                continue
            assert vis[k] == (not (1 < hit[1] < 3 and 1 < hit[2] < 3)), k


# ---------------------------------------------------------------------------
# infinite diffuse plane (Svensson & Savioja 2024), reference
# tests/test_DRadiosityFast_infinite_diffuse_plane.py: order 0 on one 20 x 20 m ground
# polygon, 1 m patches, 1 s ETC in ONE bin; ratio of diffuse to specular energy
# ---------------------------------------------------------------------------
def plane_ratio_oracle(oracle, source, receiver):
    half = 10.0
    wall = rect([-half, -half, 0], [2 * half, 0, 0], [0, 2 * half, 0])[None]
    one = np.array([[[0.0, 0.0, 1.0]]])
    out = oracle.pipeline(wall, np.array([[0.0, 0.0, 1.0]]), 1.0, np.asarray(source, float),
                          np.asarray(receiver, float)[None], 343.0, 1.0, 1.0, 0, np.zeros(1),
                          one, one, np.ones((1, 1, 1, 1)), np.zeros(1, np.int64),
                          brdf_set_before_bake=True)
    return float(out["etc_receiver_mono"].sum())


def plane_ratio_cuda(source, receiver):
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf
    half = 10.0
    plane = sp.Polygon(rect([-half, -half, 0], [2 * half, 0, 0], [0, 2 * half, 0]),
                       [1, 0, 0], [0, 0, 1])
    rad = sp.DirectionalRadiosityFast.from_polygon([plane], 1.0)
    dirs = pf.Coordinates(0, 0, 1, weights=1)
    brdf = sp.brdf.create_from_scattering(dirs, dirs, pf.FrequencyData(1, [100]),
                                          pf.FrequencyData(0, [100]))
    rad.set_wall_brdf(np.arange(1), brdf, dirs, dirs)
    rad.set_air_attenuation(pf.FrequencyData(np.zeros(1), [100]))
    rad.init_source_energy(pf.Coordinates(*source))
    rad.calculate_energy_exchange(speed_of_sound=343, etc_time_resolution=1.0,
                                  etc_duration=1, max_reflection_order=0)
    etc = rad.collect_energy_receiver_mono(pf.Coordinates(*receiver))
    return float(np.sum(etc.time))


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=[pytest.mark.gpu])])
def plane_ratio(request, oracle):
    def ratio(source, receiver):
        image = np.array([source[0], source[1], -source[2]])
        specular = 1 / (4 * np.pi * np.sum((np.asarray(receiver) - image) ** 2))
        diffuse = (plane_ratio_oracle(oracle, source, receiver) if request.param == "oracle"
                   else plane_ratio_cuda(source, receiver))
        return diffuse / specular
    return ratio


@pytest.mark.parametrize("source, receiver", [((0, 0, 3), (0, 0, 3)), ((0, 0, 5), (0, 0, 3))])
def test_diffuse_plane_along_the_normal(plane_ratio, source, receiver):
    """cases 1 and 2 of the paper: ratio 2 for an infinite plane, 1.97 for 20 x 20 m
    (reference tests/test_DRadiosityFast_infinite_diffuse_plane.py:93-131)"""
    ratio = plane_ratio(source, receiver)
    assert ratio < 2
    assert abs(ratio - 1.97) / 1.97 < 0.01


@pytest.mark.parametrize("theta_deg", [30, 45, 60])
def test_diffuse_plane_same_height(plane_ratio, theta_deg):
    """case 3: source and receiver at the same height, ratio 2 cos(theta)
    (reference tests/test_DRadiosityFast_infinite_diffuse_plane.py:134-156)"""
    th = np.deg2rad(theta_deg)
    source = (2 * np.sin(th), 0.0, 2 * np.cos(th))
    receiver = (-2 * np.sin(th), 0.0, 2 * np.cos(th))
    ratio = plane_ratio(source, receiver)
    assert abs(ratio - 2 * np.cos(th)) <= 0.03 + 0.03 * 2 * np.cos(th)


# ---------------------------------------------------------------------------
# whole-pipeline known answers: diffuse room after Kuttruff, reciprocity in a shoebox
# ---------------------------------------------------------------------------
def room_etc_oracle(oracle, dims, patch, source, receiver, c, dt, duration, order, absorption):
    walls = scenes.shoebox(*dims)
    one = np.array([[[0.0, 0.0, 1.0]]] * len(walls))
    out = oracle.pipeline(np.array([w[0] for w in walls]), np.array([w[2] for w in walls]),
                          patch, np.asarray(source, float), np.asarray(receiver, float)[None],
                          c, dt, duration, order, np.zeros(1), one, one,
                          np.full((1, 1, 1, 1), 1.0 - absorption), np.zeros(6, np.int64),
                          brdf_set_before_bake=True)
    return out["etc_receiver_mono"][0, 0]


def room_etc_cuda(dims, patch, source, receiver, c, dt, duration, order, absorption):
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf
    walls = sp.testing.shoebox_room_stub(*dims)
    rad = sp.DirectionalRadiosityFast.from_polygon(walls, patch)
    dirs = pf.Coordinates(0, 0, 1, weights=1)
    brdf = sp.brdf.create_from_scattering(dirs, dirs, pf.FrequencyData(1, [500]),
                                          pf.FrequencyData(absorption, [500]))
    rad.set_wall_brdf(np.arange(len(walls)), brdf, dirs, dirs)
    rad.set_air_attenuation(pf.FrequencyData(np.zeros(1), [500]))
    rad.bake_geometry()
    rad.init_source_energy(pf.Coordinates(*source))
    rad.calculate_energy_exchange(speed_of_sound=c, etc_time_resolution=dt,
                                  etc_duration=duration, max_reflection_order=order,
                                  recalculate=True)
    return rad.collect_energy_receiver_mono(pf.Coordinates(*receiver)).time[0, 0]


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def room_etc(request, oracle):
    if request.param == "oracle":
        return lambda *a: room_etc_oracle(oracle, *a)
    return room_etc_cuda


def test_diffuse_room_decay_after_kuttruff(room_etc):
    """reference tests/test_DRadiosityFast_order.py:49-117: perfectly diffuse 5 x 3 x 4 m
    room, absorption 0.1, 3 orders; bins 5..18 of the ETC within 2 dB of Kuttruff's
    Eq. 4.10"""
    x, y, z = 5, 3, 4
    c, dt, duration, absorption = 346.18, 1 / 500, 0.15, 0.1
    etc = room_etc((x, y, z), 1.0, (2, 1.5, 2), (3, 1.5, 2), c, dt, duration, 3, absorption)
    t = np.arange(etc.shape[-1]) * dt
    surface = 2 * (x * y + x * z + y * z)
    volume = x * y * z
    w0 = 4 / (surface * absorption) / volume
    analytic = w0 * np.exp(c * surface * np.log(1 - absorption) / (4 * volume) * (t - 0.03))
    got_db, want_db = 10 * np.log10(etc[5:19]), 10 * np.log10(analytic[5:19])
    assert np.max(np.abs(got_db - want_db)) < 2.0


@pytest.mark.parametrize("receiver", [(0.5, 1.0, 2.5), (1.0, 2.5, 1.5)])
@pytest.mark.parametrize("order", [10, 20])
@pytest.mark.parametrize("patch", [1.0, 1.5])
def test_reciprocity_in_a_shoebox(room_etc, receiver, order, patch):
    """reference tests/test_multisource.py:51-137: swapping source and receiver in a
    lossless diffuse 3 x 3 x 3 m room gives the same ETC (6 decimals there)"""
    source = (2.0, 1.5, 1.5)
    args = ((3, 3, 3), patch)
    tail = (343.0, 1 / 200, 0.5, order, 0.0)
    a = room_etc(*args, source, receiver, *tail)
    b = room_etc(*args, receiver, source, *tail)
    assert abs(a.sum() - b.sum()) < 1.5e-6
    assert np.max(np.abs(a - b)) < 1.5e-6


@pytest.mark.parametrize("src", [(2.0, 0, 0), (2.0, 2.0, 0), (2.0, 0.0, 2.0)])
@pytest.mark.parametrize("rec", [(1.0, 0.0, 0), (2.0, -2.0, 0), (2.0, 0.0, -2.0)])
def test_source_patch_receiver_reciprocity(oracle, src, rec):
    """reference tests/test_multisource.py:140-199: for one 2 x 2 m patch the product of
    the source->patch and patch->receiver point factors is symmetric in source and
    receiver (6 decimals there).  Oracle only: the CUDA point kernels are pinned to it by
    the golden point_patch vectors."""
    patch = np.array([[[0, -1, -1], [0, -1, 1], [0, 1, 1], [0, 1, -1]]], float)
    centre = patch.mean(axis=1)
    seen = np.ones(1, np.uint8)
    energy = []
    for s_, r_ in ((src, rec), (rec, src)):
        e_s, _ = oracle.source_energy(np.array(s_, float), centre, patch, seen, np.zeros(1))
        e_r = oracle.receiver_factor(np.array(r_, float), patch, seen)
        energy.append(e_s[0, 0] * e_r[0])
    assert abs(energy[0] - energy[1]) < 1.5e-6 and energy[0] > 0
