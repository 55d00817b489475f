"""Generate golden input/output vectors from the LIVE reference (sparrowpy v1.0.1).

Run once in the build container (needs /root/reference and numba):

    python tests/golden/make_golden.py

It drives the reference's own ``DirectionalRadiosityFast`` class and numba
kernels -- unmodified -- on small scenes and stores inputs + outputs as ``.npz``
fixtures next to this script.  The only thing supplied from outside the reference
is a stand-in for the absent pyfar package (``sparrowpy_b200.pyfar_shim``) and,
because ``pf.Orientations`` is part of that absent package, the per-wall BRDF
direction rotation (``_rotate_coords_to_normal``, reference
RadiosityFast.py:971-986), restated on the scipy ``Rotation`` calls that pyfar makes
(``pyfar_shim.rotate_to_wall``): the rotated direction arrays are therefore *inputs* of
the fixtures, not pinned outputs.

The fixtures pin (a) the CPU oracle in ``oracle/`` and (b) the CUDA path.
Nothing here is imported at test or bench time.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from sparrowpy_b200 import pyfar_shim, scenes  # noqa: E402

# the reference must see the shim as "pyfar" before it is imported
sys.modules["pyfar"] = pyfar_shim
from ref_import import import_reference  # noqa: E402

sp, RF, geo, ffu, integ = import_reference()
import numba  # noqa: E402


def _rotate_coords_to_normal(wall_normal, wall_up_vector, sources, receivers):
    # the reference's pyfar calls restated on scipy's Rotation (pyfar is absent here)
    return (pyfar_shim.rotate_to_wall(sources, wall_normal, wall_up_vector),
            pyfar_shim.rotate_to_wall(receivers, wall_normal, wall_up_vector))


RF._rotate_coords_to_normal = _rotate_coords_to_normal


def polygons(walls):
    return [sp.geometry.Polygon(p, u, n) for (p, u, n) in walls]


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"  wrote {name}.npz  {os.path.getsize(path)/1024:.0f} KiB")


# ---------------------------------------------------------------------------
def run_scene(name, walls, patch_size, source, receivers, c, dt, duration,
              max_order, brdf_sets=None, directions=None, air=None,
              frequencies=None, etc_rows=None, store_tilde_rows=None):
    """Full pipeline through the reference class.

    brdf_sets: list of (wall_indexes, brdf (S,D,B) not yet *pi); None = the
    class's own default 1x1 unit BRDF (installed by init_source_energy).
    """
    t0 = time.time()
    polys = polygons(walls)
    rad = sp.DirectionalRadiosityFast.from_polygon(polys, patch_size)
    out = dict(
        walls_points=np.array([w[0] for w in walls]),
        walls_up=np.array([w[1] for w in walls]),
        walls_normal=np.array([w[2] for w in walls]),
        patch_size=np.float64(patch_size), source=np.asarray(source, float),
        receivers=np.atleast_2d(np.asarray(receivers, float)),
        speed_of_sound=np.float64(c), dt=np.float64(dt),
        duration=np.float64(duration), max_order=np.int64(max_order),
    )
    if brdf_sets is not None:
        dirs, weights = directions
        coords = pyfar_shim.Coordinates.from_cartesian(dirs, weights=weights)
        for wall_idx, brdf in brdf_sets:
            rad.set_wall_brdf(
                np.asarray(wall_idx), pyfar_shim.FrequencyData(brdf, frequencies),
                coords, coords)
        out["brdf_dirs"] = dirs
        out["brdf_weights"] = weights
        out["frequencies"] = np.asarray(frequencies, float)
    if air is not None:
        rad.set_air_attenuation(pyfar_shim.FrequencyData(air, frequencies))

    out["patches_points"] = rad.patches_points
    out["patch_to_wall_ids"] = rad._patch_to_wall_ids
    out["patches_center"] = rad.patches_center
    out["patches_area"] = rad.patches_area
    out["patches_normal"] = rad.patches_normal

    rad.bake_geometry()
    n = rad.n_patches
    out["visibility"] = np.packbits(rad._visibility_matrix, axis=None)
    out["visible_patches"] = rad._visible_patches
    vp = rad._visible_patches
    out["ff_pairs"] = rad._form_factors[vp[:, 0], vp[:, 1]]
    out["ff_nnz_outside_pairs"] = np.int64(
        np.count_nonzero(rad._form_factors) - np.count_nonzero(out["ff_pairs"]))
    out["p2o"] = rad._patch_2_brdf_outgoing_index.astype(np.int16)

    src = pyfar_shim.Coordinates(*np.asarray(source, float))
    rad.init_source_energy(src)
    # (after this call the class has installed default BRDF / air if unset)
    out["air_attenuation"] = np.asarray(rad._air_attenuation, float)
    out["brdf"] = np.real(np.array(rad._brdf)).reshape(
        (len(rad._brdf),) + tuple(np.array(rad._brdf).shape[1:]))
    out["brdf_index"] = np.asarray(rad._brdf_index, np.int64)
    out["vi"] = np.array([s.cartesian for s in rad._brdf_incoming_directions])
    out["vo"] = np.array([s.cartesian for s in rad._brdf_outgoing_directions])
    out["source_visibility"] = rad._source_visibility
    out["energy_init_source"] = rad._energy_init_source
    out["distance_patches_to_source"] = rad._distance_patches_to_source
    e0, _ = ffu._source2patch_energy_universal(
        np.asarray(source, float), rad.patches_center, rad.patches_points,
        rad._source_visibility, rad._air_attenuation, rad.n_bins)
    out["energy_0"] = e0

    # form_factors_tilde: when the BRDF was only installed by init_source_energy
    # the baked tilde has shape (N,N,1,1)
    tilde = rad._form_factors_tilde
    if store_tilde_rows is None:
        out["tilde"] = tilde
    else:
        out["tilde_rows"] = np.asarray(store_tilde_rows, np.int64)
        out["tilde_sample"] = tilde[np.asarray(store_tilde_rows)]

    rad.calculate_energy_exchange(c, dt, duration, max_reflection_order=max_order)
    etc = rad._energy_exchange_etc
    out["etc_shape"] = np.array(etc.shape, np.int64)
    if etc_rows is None:
        out["etc"] = etc
    else:
        out["etc_rows"] = np.asarray(etc_rows, np.int64)
        out["etc_sample"] = etc[np.asarray(etc_rows)]
    out["etc_patch_sums"] = etc.sum(axis=-1)

    # order-0 only variant (reference RadiosityFast.py:550-555)
    rad.calculate_energy_exchange(c, dt, duration, max_reflection_order=0,
                                  recalculate=True)
    out["etc_order0_sums"] = rad._energy_exchange_etc.sum(axis=-1)
    rad.calculate_energy_exchange(c, dt, duration, max_reflection_order=max_order,
                                  recalculate=True)

    # pair delays exactly as the reference computes them (RadiosityFast.py:538-543,
    # :1135-1136)
    cen = rad.patches_center
    delays = np.empty(len(vp), np.int64)
    for k, (i, j) in enumerate(vp):
        d = np.linalg.norm(cen[i, :] - cen[j, :])
        delays[k] = int(d / c / dt)
    out["pair_delays"] = delays
    out["source_delays"] = np.array(
        [int(d / c / dt) for d in rad._distance_patches_to_source], np.int64)

    rec = np.atleast_2d(np.asarray(receivers, float))
    rcoords = pyfar_shim.Coordinates.from_cartesian(rec)
    patchwise = rad.collect_energy_receiver_patchwise(rcoords).time
    mono = rad.collect_energy_receiver_mono(rcoords).time
    out["etc_receiver_mono"] = mono
    out["etc_receiver_patch_sums"] = patchwise.sum(axis=-1)
    rvis, rfac, ridx, rdel = [], [], [], []
    for r in rec:
        v = geo._check_point2patch_visibility(
            eval_point=r, patches_center=cen, surf_points=rad.walls_points,
            surf_normal=rad.walls_normal)
        rvis.append(v)
        rfac.append(ffu._patch2receiver_energy_universal(
            r, rad.patches_points, v))
        ridx.append(RF.get_scattering_data_receiver_index(
            cen, r, out["vo"], rad._patch_to_wall_ids))
        dist = np.linalg.norm(cen - r, axis=1)
        rdel.append(np.array([int(np.ceil(d / c / dt)) for d in dist]))
    out["receiver_visibility"] = np.array(rvis)
    out["receiver_factor"] = np.array(rfac)
    out["receiver_dir_index"] = np.array(ridx, np.int64)
    out["receiver_delays"] = np.array(rdel, np.int64)
    save(name, **out)
    print(f"  {name}: N={n} P={len(vp)} etc={etc.shape} "
          f"({time.time()-t0:.1f}s)")


# ---------------------------------------------------------------------------
def gen_tessellation():
    cases = {}
    specs = [
        ("cube05", scenes.shoebox(1, 1, 1), 0.5),
        ("box564_1", scenes.shoebox(5, 6, 4), 1.0),
        ("box322_02", scenes.shoebox(3, 2, 2), 0.2),   # int(0.6/0.2) hazards
        ("box111_03", scenes.shoebox(1, 1, 1), 0.3),
        ("canyon01", scenes.street_canyon(seed=0, scale=0.1), 1.0),
    ]
    for name, walls, ps in specs:
        wp = np.array([w[0] for w in walls])
        wn = np.array([w[2] for w in walls])
        pts, nrm, n, ids = geo._process_patches(wp, wn, ps, len(walls))
        cases[name + "_walls"] = wp
        cases[name + "_normals"] = wn
        cases[name + "_size"] = np.float64(ps)
        cases[name + "_points"] = pts
        cases[name + "_ids"] = ids
        cases[name + "_center"] = geo._calculate_center(pts)
        cases[name + "_area"] = geo._calculate_area(pts)
    save("tessellation", **cases)


def _random_rect(rng, lattice=None):
    """Planar rectangle with a random orientation (general normal)."""
    a = rng.normal(size=3)
    a /= np.linalg.norm(a)
    b = rng.normal(size=3)
    b -= np.dot(a, b) * a
    b /= np.linalg.norm(b)
    n = np.cross(a, b)
    o = rng.uniform(-2, 2, size=3)
    la, lb = rng.uniform(0.3, 2.0, size=2)
    pts = np.array([o, o + la * a, o + la * a + lb * b, o + lb * b])
    return pts, n


def _axis_rect(rng, h):
    axis = int(rng.integers(0, 3))
    others = [k for k in range(3) if k != axis]
    o = rng.integers(-4, 5, size=3) * h
    la, lb = rng.integers(1, 4, size=2) * h
    pts = np.tile(o.astype(float), (4, 1))
    pts[1, others[0]] += la
    pts[2, others[0]] += la
    pts[2, others[1]] += lb
    pts[3, others[1]] += lb
    n = np.zeros(3)
    n[axis] = rng.choice([-1.0, 1.0])
    return pts, n


def gen_predicates():
    rng = np.random.default_rng(1234)
    A, B, S, Nn = [], [], [], []
    # random general-position cases
    for _ in range(1500):
        pts, n = _random_rect(rng)
        A.append(rng.uniform(-3, 3, 3))
        B.append(rng.uniform(-3, 3, 3))
        S.append(pts)
        Nn.append(n)
    # random rectangles with an endpoint inside / on the plane
    for _ in range(600):
        pts, n = _random_rect(rng)
        u, v = rng.uniform(-0.2, 1.2, 2)
        p_in = pts[0] + u * (pts[1] - pts[0]) + v * (pts[3] - pts[0])
        other = rng.uniform(-3, 3, 3)
        if rng.random() < 0.3:
            u, v = rng.uniform(-0.2, 1.2, 2)
            other = pts[0] + u * (pts[1] - pts[0]) + v * (pts[3] - pts[0])
        if rng.random() < 0.5:
            A.append(p_in), B.append(other)
        else:
            A.append(other), B.append(p_in)
        S.append(pts)
        Nn.append(n)
    # lattice-degenerate axis-aligned cases (segments through edges/vertices)
    for h in (1.0, 0.5, 0.2, 1.0 / 3.0):
        for _ in range(1200):
            pts, n = _axis_rect(rng, h)
            A.append((rng.integers(-8, 9, 3) * 0.5) * h)
            B.append((rng.integers(-8, 9, 3) * 0.5) * h)
            S.append(pts)
            Nn.append(n)
    A, B, S, Nn = map(np.array, (A, B, S, Nn))
    vis = np.array([geo._basic_visibility(A[k], B[k], S[k], Nn[k])
                    for k in range(len(A))])
    inA = np.array([geo._point_in_polygon(A[k], S[k], Nn[k])
                    for k in range(len(A))])
    inB = np.array([geo._point_in_polygon(B[k], S[k], Nn[k])
                    for k in range(len(A))])
    print(f"  predicates: {len(A)} cases, visible={vis.sum()}, inA={inA.sum()}, "
          f"inB={inB.sum()}")
    save("predicates", A=A, B=B, S=S, N=Nn, visible=vis, inA=inA, inB=inB)


def gen_rounding_probes():
    """Pin the arithmetic model of numba's np.dot / np.linalg.norm."""
    rng = np.random.default_rng(99)

    @numba.njit()
    def _norms(v):
        out = np.empty(v.shape[0])
        for k in range(v.shape[0]):
            out[k] = np.linalg.norm(v[k])
        return out

    @numba.njit()
    def _dots(a, b):
        out = np.empty(a.shape[0])
        for k in range(a.shape[0]):
            out[k] = np.dot(a[k], b[k])
        return out

    v3 = rng.normal(size=(4000, 3)) * 10.0 ** rng.integers(-3, 3, (4000, 1))
    v2 = rng.normal(size=(4000, 2)) * 10.0 ** rng.integers(-3, 3, (4000, 1))
    # lattice differences (typical patch-centre differences)
    l3 = rng.integers(-400, 401, (4000, 3)) * 0.05
    a3, b3 = rng.normal(size=(2, 4000, 3))
    a2, b2 = rng.normal(size=(2, 4000, 2))
    np_norm1d = np.array([np.linalg.norm(x) for x in l3])
    save("rounding_probes",
         v3=v3, n3=_norms(v3), v2=v2, n2=_norms(v2), l3=l3, nl3=_norms(l3),
         a3=a3, b3=b3, d3=_dots(a3, b3), a2=a2, b2=b2, d2=_dots(a2, b2),
         np_norm1d_l3=np_norm1d, np_norm_axis1_l3=np.linalg.norm(l3, axis=1))


def gen_form_factor_pairs():
    """Patch pairs straight through universal_form_factor (universal.py:54-96)."""
    rng = np.random.default_rng(7)
    PI, NI, AI, PJ, NJ = [], [], [], [], []

    def add(pi, ni, pj, nj):
        PI.append(pi), NI.append(ni), PJ.append(pj), NJ.append(nj)
        AI.append(geo._polygon_area(np.asarray(pi, float)))

    def rect(axis, const, lo, hi, sign):
        p, _, n = scenes._rect(axis, const, lo, hi, sign)
        return p, n

    # parallel facing squares at several distances / sizes (Stokes)
    for w, d in [(1, 1), (1, 0.5), (2, 1), (0.5, 3), (1, 10), (0.2, 40)]:
        pi, ni = rect(2, 0.0, (0, 0), (w, w), 1.0)
        pj, nj = rect(2, float(d), (0, 0), (w, w), -1.0)
        add(pi, ni, pj, nj)
    # perpendicular sharing an edge (Nusselt branch), equal and unequal sizes
    for w, l, hgt in [(1, 1, 1), (2, 1, 1), (1, 2, 0.5), (0.5, 0.5, 0.5)]:
        pi, ni = rect(2, 0.0, (0, 0), (w, l), 1.0)
        pj, nj = rect(1, 0.0, (0, 0), (w, hgt), 1.0)
        add(pi, ni, pj, nj)
    # perpendicular sharing only a vertex (Nusselt branch)
    for w in (1.0, 0.5):
        pi, ni = rect(2, 0.0, (0, 0), (w, w), 1.0)
        pj, nj = rect(1, 0.0, (w, 0), (2 * w, w), 1.0)
        add(pi, ni, pj, nj)
    # perpendicular, offset (no shared vertex -> Stokes)
    pi, ni = rect(2, 0.0, (0, 1), (1, 2), 1.0)
    pj, nj = rect(1, 0.0, (0, 0), (1, 1), 1.0)
    add(pi, ni, pj, nj)
    # random patch pairs from a tessellated shoebox (both branches occur)
    walls = scenes.shoebox(3, 4, 2)
    wp = np.array([w[0] for w in walls])
    wn = np.array([w[2] for w in walls])
    pts, nrm, n, ids = geo._process_patches(wp, wn, 0.5, len(walls))
    for _ in range(150):
        i, j = rng.integers(0, n, 2)
        if ids[i] == ids[j]:
            continue
        add(pts[i], nrm[i], pts[j], nrm[j])
    PI, NI, AI, PJ, NJ = map(lambda x: np.array(x, float), (PI, NI, AI, PJ, NJ))
    ff = np.array([ffu.universal_form_factor(PI[k], NI[k], AI[k], PJ[k], NJ[k])
                   for k in range(len(PI))])
    nus = np.array([geo._coincidence_check(PJ[k], PI[k]) for k in range(len(PI))])
    print(f"  ff pairs: {len(ff)} ({nus.sum()} Nusselt)")
    save("form_factor_pairs", pts_i=PI, normal_i=NI, area_i=AI, pts_j=PJ,
         normal_j=NJ, ff=ff, nusselt=nus)


def gen_point_patch():
    rng = np.random.default_rng(21)
    P, Q = [], []
    for _ in range(300):
        axis = int(rng.integers(0, 3))
        lo = rng.uniform(-2, 2, 2)
        hi = lo + rng.uniform(0.2, 2, 2)
        pts, _, _ = scenes._rect(axis, float(rng.uniform(-2, 2)), lo, hi, 1.0)
        P.append(pts)
        Q.append(rng.uniform(-4, 4, 3))
    P, Q = np.array(P), np.array(Q)
    src = np.array([integ.pt_solution(Q[k], P[k], mode="source")
                    for k in range(len(P))])
    rcv = np.array([integ.pt_solution(Q[k], P[k], mode="receiver")
                    for k in range(len(P))])
    save("point_patch", patches=P, points=Q, source=src, receiver=rcv)


def gen_directivity_metrics():
    """Azimuth / elevation of a target in the source frame: the reference's
    sound_object._get_metrics (sound_object.py:67-87), pure numpy."""
    from sparrowpy import sound_object as so
    rng = np.random.default_rng(42)
    n = 300
    pos, tgt = rng.uniform(-3, 3, (n, 3)), rng.uniform(-6, 6, (n, 3))
    view = rng.normal(size=(n, 3))
    view /= np.linalg.norm(view, axis=1)[:, None]
    up = rng.normal(size=(n, 3))
    up -= np.sum(up * view, 1)[:, None] * view
    up /= np.linalg.norm(up, axis=1)[:, None]
    # axis-aligned frames with targets on the axes (branch cuts of arctan2 / arcsin)
    view[:6] = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [1, 0, 0]]
    up[:6] = [[0, 0, 1], [0, 0, 1], [1, 0, 0], [0, 0, 1], [0, 0, 1], [0, 1, 0]]
    tgt[:6] = pos[:6] + np.array([[1, 0, 0], [0, 0, 2], [0, 1, 0], [1, 1, 0], [0, 0, -1],
                                  [0, -3, 0]])
    az, el = np.zeros(n), np.zeros(n)
    for k in range(n):
        az[k], el[k] = so._get_metrics(pos[k], view[k], up[k], tgt[k])
    save("directivity_metrics", pos=pos, view=view, up=up, target=tgt, azimuth_deg=az,
         elevation_deg=el)


def gen_brdf_builders():
    """The reference's BRDF builders (sparrowpy/brdf.py:8-224) driven with the pyfar
    stand-in: direction sets with unnormalised weights, several bands, with and without
    absorption.  Pins sparrowpy_b200/brdf.py."""
    from sparrowpy import brdf as ref_brdf
    rng = np.random.default_rng(7)
    out = {}
    for tag, (n_az, cols) in {"a": (4, (45.0,)), "b": (8, (30.0, 60.0)),
                              "c": (6, (20.0, 50.0, 75.0))}.items():
        dirs, w = scenes.hemisphere_directions(n_az, cols)
        w = w * rng.uniform(0.5, 2.0)                       # weights need not be normalised
        freqs = np.array([125.0, 500.0, 2000.0])
        s = rng.uniform(0.05, 0.95, 3)
        alpha = rng.uniform(0.0, 0.6, 3)
        src = pyfar_shim.Coordinates.from_cartesian(dirs, weights=w.copy())
        rcv = pyfar_shim.Coordinates.from_cartesian(dirs, weights=w.copy())
        b1 = ref_brdf.create_from_scattering(
            src, rcv, pyfar_shim.FrequencyData(s, freqs), pyfar_shim.FrequencyData(alpha, freqs))
        rcv2 = pyfar_shim.Coordinates.from_cartesian(dirs, weights=w.copy())
        b2 = ref_brdf.create_from_scattering(src, rcv2, pyfar_shim.FrequencyData(s, freqs))
        sd = rng.uniform(0.0, 1.0, (len(dirs), len(dirs), 3))
        sd /= sd.sum(axis=1, keepdims=True)
        rcv3 = pyfar_shim.Coordinates.from_cartesian(dirs, weights=w.copy())
        b3 = ref_brdf.create_from_directional_scattering(
            src, rcv3, pyfar_shim.FrequencyData(sd, freqs),
            pyfar_shim.FrequencyData(alpha, freqs))
        out.update({f"{tag}_dirs": dirs, f"{tag}_weights": w, f"{tag}_freqs": freqs,
                    f"{tag}_scattering": s, f"{tag}_absorption": alpha,
                    f"{tag}_directional": sd,
                    f"{tag}_brdf_scattering": np.real(b1.freq),
                    f"{tag}_brdf_scattering_noabs": np.real(b2.freq),
                    f"{tag}_brdf_directional": np.real(b3.freq),
                    f"{tag}_weights_after": np.asarray(rcv.weights)})
    save("brdf_builders", **out)


def gen_direct_sound():
    """``calculate_direct_sound`` (RadiosityFast.py:605-657) of the reference class, called
    unbound on a minimal stand-in object (it only reads the attributes set below)."""
    import types
    rng = np.random.default_rng(11)
    src = np.array([1.3, -0.4, 1.7])
    rcv = rng.uniform(-6, 6, (12, 3))
    air = np.array([0.0, 1e-3, 7e-3])
    c, dt = 343.2, 0.5e-3
    obj = types.SimpleNamespace(
        _source=pyfar_shim.Coordinates(*src), n_bins=3, _air_attenuation=air,
        speed_of_sound=c, _etc_time_resolution=dt, _frequencies=np.array([250., 1e3, 4e3]))
    direct, delay = RF.DirectionalRadiosityFast.calculate_direct_sound(
        obj, pyfar_shim.Coordinates.from_cartesian(rcv))
    save("direct_sound", source=src, receivers=rcv, air_attenuation=air, speed_of_sound=c,
         dt=dt, direct_sound=direct, n_sample_delay=delay)


def gen_wall_rotation():
    """Invariants of ``_rotate_coords_to_normal`` that the reference's tests state
    (tests/test_DRadiosityFast.py:46-107, :153-172) evaluated on the restated rotation, plus
    the rotated direction sets themselves (regression vectors for the shim)."""
    d, w = scenes.hemisphere_directions(8, (30.0, 60.0))
    coords = pyfar_shim.Coordinates.from_cartesian(d, weights=w)
    frames = [([1, 0, 0], [0, 0, 1]), ([0, 1, 0], [0, 0, -1]), ([0, 0, 1], [1, 0, 0]),
              ([0, 0, -1], [0, 1, 0]), ([-1, 0, 0], [0, 1, 0]), ([0, -1, 0], [1, 0, 0])]
    rng = np.random.default_rng(5)
    for _ in range(6):
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        u = rng.normal(size=3)
        u -= np.dot(u, n) * n
        u /= np.linalg.norm(u)
        frames.append((n, u))
    normals = np.array([f[0] for f in frames], float)
    ups = np.array([f[1] for f in frames], float)
    rotated = np.array([_rotate_coords_to_normal(n, u, coords, coords)[0].cartesian
                        for n, u in zip(normals, ups)])
    save("wall_rotation", dirs=d, weights=w, normals=normals, ups=ups, rotated=rotated)


def main():
    only = set(sys.argv[1:])

    def want(tag):
        return not only or tag in only

    if want("tess"):
        print("tessellation"); gen_tessellation()
    if want("round"):
        print("rounding probes"); gen_rounding_probes()
    if want("pred"):
        print("predicates"); gen_predicates()
    if want("ffpairs"):
        print("form factor pairs"); gen_form_factor_pairs()
    if want("ptpatch"):
        print("point-patch"); gen_point_patch()
    if want("directivity"):
        print("directivity metrics"); gen_directivity_metrics()
    if want("brdf"):
        print("brdf builders"); gen_brdf_builders()
    if want("direct"):
        print("direct sound"); gen_direct_sound()
    if want("rotation"):
        print("wall rotation"); gen_wall_rotation()

    if want("cube"):
        print("scene cube05 (reference tests/test_DRadiosityFast.py:19-29)")
        run_scene("scene_cube05", scenes.shoebox(1, 1, 1), 0.5,
                  source=[0.5, 0.5, 0.5], receivers=[[0.25, 0.5, 0.5]],
                  c=343.2, dt=0.5e-3, duration=0.03, max_order=5)
    if want("c1"):
        print("scene C1 (BASELINE config 1)")
        dirs = (np.array([[0.0, 0.0, 1.0]]), np.array([1.0]))
        run_scene("scene_c1", scenes.shoebox(5, 6, 4), 1.0,
                  source=[2, 2, 2], receivers=[[2, 3, 2]],
                  c=343.2, dt=1e-3, duration=1.0, max_order=150,
                  brdf_sets=[(np.arange(6), np.full((1, 1, 1), 0.9 / np.pi))],
                  directions=dirs, air=np.zeros(1), frequencies=[1000.0],
                  etc_rows=np.arange(0, 148, 9), store_tilde_rows=[0, 37, 147])
    if want("occ"):
        print("scene occluder")
        run_scene("scene_occluder", scenes.occluder_scene(6, 2, 2), 1.0,
                  source=[1.0, 1.2, 1.5], receivers=[[5.0, 4.6, 1.2],
                                                     [1.0, 5.0, 0.5]],
                  c=343.2, dt=0.5e-3, duration=0.06, max_order=6)
    if want("dir"):
        print("scene directional")
        d4, w4 = scenes.hemisphere_directions(4, (45.0,))
        freqs = [500.0, 2000.0]
        b0 = scenes.brdf_from_scattering(d4, w4, [0.5, 0.7], [0.1, 0.2])
        b1 = scenes.brdf_from_scattering(d4, w4, [1.0, 1.0], [0.3, 0.05])
        run_scene("scene_directional", scenes.shoebox(3, 2, 2), 0.5,
                  source=[1.1, 0.9, 1.2], receivers=[[2.2, 1.3, 0.7],
                                                     [0.4, 0.5, 1.6]],
                  c=343.2, dt=0.25e-3, duration=0.03, max_order=4,
                  brdf_sets=[([0, 1, 2, 3], b0), ([4, 5], b1)],
                  directions=(d4, w4), air=np.array([1e-3, 4e-3]),
                  frequencies=freqs, etc_rows=np.arange(0, 128, 8),
                  store_tilde_rows=[0, 5, 50, 101, 127])
    if want("uneven"):
        # patch size that does not divide the walls: int(size / patch) truncation
        # (geometry.py:330-345) gives 0.75 x 0.7 x 0.7 m patches; two BRDF sets, air
        print("scene uneven (non-dividing patch size, two BRDF sets)")
        d4, w4 = scenes.hemisphere_directions(4, (45.0,))
        b0 = scenes.brdf_from_scattering(d4, w4, [0.3, 0.9], [0.15, 0.1])
        b1 = scenes.brdf_from_scattering(d4, w4, [0.8, 0.2], [0.05, 0.3])
        run_scene("scene_uneven", scenes.shoebox(3, 2.1, 1.4), 0.7,
                  source=[0.9, 1.3, 0.6], receivers=[[2.4, 0.5, 1.0], [0.3, 1.8, 0.2]],
                  c=343.2, dt=0.25e-3, duration=0.025, max_order=3,
                  brdf_sets=[([0, 2, 4], b0), ([1, 3, 5], b1)],
                  directions=(d4, w4), air=np.array([2e-3, 9e-3]),
                  frequencies=[250.0, 4000.0])
    if want("canyondir"):
        print("scene canyon (scale 0.15) with a directional BRDF")
        d4, w4 = scenes.hemisphere_directions(4, (45.0,))
        b0 = scenes.brdf_from_scattering(d4, w4, [0.4, 0.8], [0.2, 0.1])
        walls = scenes.street_canyon(seed=1, scale=0.15)
        run_scene("scene_canyon015_dir", walls, 1.0,
                  source=[2.5, 4.5, 1.5], receivers=[[15.5, 4.5, 1.5], [9.5, 3.5, 2.5]],
                  c=343.2, dt=0.5e-3, duration=0.1, max_order=3,
                  brdf_sets=[(np.arange(len(walls)), b0)], directions=(d4, w4),
                  air=np.array([1e-3, 5e-3]), frequencies=[500.0, 2000.0],
                  etc_rows=np.arange(0, 400, 25), store_tilde_rows=[0, 100, 399])
    if want("canyon"):
        print("scene canyon (scale 0.1)")
        run_scene("scene_canyon01", scenes.street_canyon(seed=0, scale=0.1), 1.0,
                  source=[1.5, 0.5, 1.5], receivers=[[10.5, 5.5, 1.5]],
                  c=343.2, dt=0.5e-3, duration=0.08, max_order=4,
                  etc_rows=np.arange(0, 192, 12),
                  store_tilde_rows=[0, 80, 191])


if __name__ == "__main__":
    main()
