"""ctypes front-end of the CPU oracle (``sparrow_oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and
bench.py's ``cpu_baseline`` / ``--impl reference`` leg -- never by the product
package ``sparrowpy_b200``.  Parity status: pinned against live-reference vectors
(tests/test_oracle_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsparrow_oracle.so")
_lib = None

c_double_p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile the oracle with gcc (see oracle/Makefile)."""
    src = os.path.join(_HERE, "sparrow_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        for name in ("sor_dot3", "sor_dot2", "sor_nrm3", "sor_nrm2",
                     "sor_np_norm1d3", "sor_np_norm_axis3",
                     "sor_universal_form_factor", "sor_pt_solution"):
            getattr(_lib, name).restype = ctypes.c_double
        _lib.sor_count_patches.restype = ctypes.c_int64
        _lib.sor_nearest_direction.restype = ctypes.c_int64
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _d(x):
    return ctypes.c_double(float(x))


def _l(x):
    return ctypes.c_int64(int(x))


# -- primitives ---------------------------------------------------------------
def vec_apply(fn_name, *arrays):
    """Apply a scalar probe function row-wise (for the rounding probes)."""
    fn = getattr(lib(), fn_name)
    arrays = [_f64(a) for a in arrays]
    n = arrays[0].shape[0]
    out = np.empty(n)
    for k in range(n):
        out[k] = fn(*[_p(a[k]) for a in arrays])
    return out


# -- tessellation ---------------------------------------------------------------
def process_patches(walls_points, patch_size):
    walls = _f64(walls_points)
    n = lib().sor_count_patches(_p(walls), _l(walls.shape[0]), _d(patch_size))
    pts = np.empty((n, 4, 3))
    ids = np.empty(n, np.int64)
    lib().sor_create_patches(_p(walls), _l(walls.shape[0]), _d(patch_size),
                             _p(pts), _p(ids))
    return pts, ids


def centers(points):
    points = _f64(points)
    out = np.empty((points.shape[0], 3))
    lib().sor_centers(_p(points), _l(points.shape[0]), _p(out))
    return out


def areas(points):
    points = _f64(points)
    out = np.empty(points.shape[0])
    lib().sor_areas(_p(points), _l(points.shape[0]), _p(out))
    return out


# -- visibility -------------------------------------------------------------------
def basic_visibility(a, b, surf, normal):
    surf = _f64(surf)
    return bool(lib().sor_basic_visibility(
        _p(_f64(a)), _p(_f64(b)), _p(surf), ctypes.c_int(surf.shape[0]),
        _p(_f64(normal))))


def point_in_polygon(p, surf, normal):
    surf = _f64(surf)
    return bool(lib().sor_point_in_polygon(
        _p(_f64(p)), _p(surf), ctypes.c_int(surf.shape[0]), _p(_f64(normal))))


def visibility_p2p(centers_, surf_normals, surf_points, row_lo=0, row_hi=None):
    c, sn, sp = _f64(centers_), _f64(surf_normals), _f64(surf_points)
    n, m, nv = c.shape[0], sn.shape[0], sp.shape[1]
    row_hi = n if row_hi is None else row_hi
    vis = np.zeros((row_hi - row_lo, n), np.uint8)
    lib().sor_visibility_p2p_rows(_p(c), _p(sn), _p(sp), _l(n), _l(m),
                                  ctypes.c_int(nv), _l(row_lo), _l(row_hi), _p(vis))
    return vis.astype(bool)


def visibility_pt2p(point, centers_, surf_normals, surf_points):
    c, sn, sp = _f64(centers_), _f64(surf_normals), _f64(surf_points)
    vis = np.zeros(c.shape[0], np.uint8)
    lib().sor_visibility_pt2p(_p(_f64(point)), _p(c), _p(sn), _p(sp),
                              _l(c.shape[0]), _l(sn.shape[0]),
                              ctypes.c_int(sp.shape[1]), _p(vis))
    return vis.astype(bool)


def visible_pairs(vis):
    """Row-major list of True entries, int32 (reference RadiosityFast.py:377-387)."""
    i, j = np.nonzero(vis)
    return np.stack([i, j], axis=1).astype(np.int32)


# -- form factors -----------------------------------------------------------------
def universal_form_factor(pts_i, n_i, area_i, pts_j, n_j):
    return lib().sor_universal_form_factor(
        _p(_f64(pts_i)), _p(_f64(n_i)), _d(area_i), _p(_f64(pts_j)), _p(_f64(n_j)))


def coincidence_check(p0, p1):
    return bool(lib().sor_coincidence_check(_p(_f64(p0)), _p(_f64(p1))))


def ff_pairs(points, normals, areas_, pairs):
    pairs = np.ascontiguousarray(pairs, np.int32)
    out = np.empty(pairs.shape[0])
    lib().sor_ff_pairs(_p(_f64(points)), _p(_f64(normals)), _p(_f64(areas_)),
                       _p(pairs), _l(pairs.shape[0]), _p(out))
    return out


def pt_solution(point, patch, mode):
    return lib().sor_pt_solution(_p(_f64(point)), _p(_f64(patch)),
                                 ctypes.c_int(0 if mode == "source" else 1))


def source_energy(src, centers_, points, vis, air):
    c = _f64(centers_)
    air = _f64(air)
    n, nb = c.shape[0], air.shape[0]
    energy = np.empty((n, nb))
    dist = np.empty(n)
    lib().sor_source_energy(_p(_f64(src)), _p(c), _p(_f64(points)),
                            _p(np.ascontiguousarray(vis, np.uint8)), _p(air),
                            _l(n), _l(nb), _p(energy), _p(dist))
    return energy, dist


def receiver_factor(rcv, points, vis):
    points = _f64(points)
    out = np.empty(points.shape[0])
    lib().sor_receiver_factor(_p(_f64(rcv)), _p(points),
                              _p(np.ascontiguousarray(vis, np.uint8)),
                              _l(points.shape[0]), _p(out))
    return out


def receiver_dir_index(pos_i, pos_j, vo, wall_id):
    pos_i, vo = _f64(pos_i), _f64(vo)
    out = np.empty(pos_i.shape[0], np.int64)
    lib().sor_receiver_dir_index(_p(pos_i), _l(pos_i.shape[0]), _p(_f64(pos_j)),
                                 _p(vo), _l(vo.shape[1]), _p(_i64(wall_id)), _p(out))
    return out


def add_directional(energy0, src, centers_, patch_to_wall, vi, vo, brdf, brdf_index):
    e0, vi, brdf = _f64(energy0), _f64(vi), _f64(brdf)
    n, nb = e0.shape
    n_out = np.asarray(vo).shape[1]
    out = np.empty((n, n_out, nb))
    lib().sor_add_directional(
        _p(e0), _p(_f64(src)), _p(_f64(centers_)), _p(_i64(patch_to_wall)), _p(vi),
        _l(vi.shape[1]), _l(n_out), _p(brdf), _p(_i64(brdf_index)), _l(n), _l(nb),
        _p(out))
    return out


def pair_tables(centers_, areas_, patch_to_wall, pairs, ff, air, vi, vo, brdf,
                brdf_index, c, dt):
    pairs = np.ascontiguousarray(pairs, np.int32)
    vi, vo, brdf, air = _f64(vi), _f64(vo), _f64(brdf), _f64(air)
    npairs, nb = pairs.shape[0], air.shape[0]
    n_in, n_out = vi.shape[1], vo.shape[1]
    tilde = np.empty((2 * npairs, n_out, nb))
    out_dir = np.empty(2 * npairs, np.int64)
    in_dir = np.empty(2 * npairs, np.int64)
    delay = np.empty(2 * npairs, np.int64)
    lib().sor_pair_tables(
        _p(_f64(centers_)), _p(_f64(areas_)), _p(_i64(patch_to_wall)), _p(pairs),
        _p(_f64(ff)), _l(npairs), _p(air), _p(vi), _l(n_in), _p(vo), _l(n_out),
        _p(brdf), _p(_i64(brdf_index)), _l(nb), _d(c), _d(dt), _p(tilde),
        _p(out_dir), _p(in_dir), _p(delay))
    return tilde, out_dir, in_dir, delay


def init_energy(e0, distance0, n_samples, c, dt):
    e0 = _f64(e0)
    n, nd, nb = e0.shape
    etc = np.empty((n, nd, nb, n_samples))
    lib().sor_init_energy(_p(e0), _p(_f64(distance0)), _l(n), _l(nd), _l(nb),
                          _l(n_samples), _d(c), _d(dt), _p(etc))
    return etc


def energy_exchange(e0, distance0, pairs, tilde, out_dir, delay, n_samples, c, dt,
                    max_order, n_threads=1):
    e0 = _f64(e0)
    pairs = np.ascontiguousarray(pairs, np.int32)
    n, nd, nb = e0.shape
    etc = np.empty((n, nd, nb, n_samples))
    work = np.empty((2, n, nd, nb, n_samples)) if max_order >= 1 else np.empty(1)
    lib().sor_energy_exchange(
        _p(e0), _p(_f64(distance0)), _p(pairs), _l(pairs.shape[0]), _p(_f64(tilde)),
        _p(_i64(out_dir)), _p(_i64(delay)), _l(n), _l(nd), _l(nb), _l(n_samples),
        _d(c), _d(dt), _l(max_order), ctypes.c_int(n_threads), _p(etc), _p(work))
    return etc


def collect_receiver(etc, rcv, centers_, factor, dir_index, air, c, dt,
                     patchwise=False):
    etc = _f64(etc)
    n, nd, nb, t = etc.shape
    mono = np.empty((nb, t))
    pw = np.zeros((n, nb, t)) if patchwise else None
    delays = np.empty(n, np.int64)
    lib().sor_collect_receiver(
        _p(etc), _p(_f64(rcv)), _p(_f64(centers_)), _p(_f64(factor)),
        _p(_i64(dir_index)), _p(_f64(air)), _l(n), _l(nd), _l(nb), _l(t), _d(c),
        _d(dt), _p(pw) if patchwise else None, _p(mono), _p(delays))
    return mono, pw, delays


def max_threads():
    return int(lib().sor_max_threads())


# -- whole pipeline on raw arrays ----------------------------------------------------
def pipeline(walls_points, walls_normal, patch_size, source, receivers, c, dt,
             duration, max_order, air, vi, vo, brdf, brdf_index, n_threads=1,
             brdf_set_before_bake=True):
    """The reference's call sequence on plain arrays.

    ``vi``/``vo``: per-wall rotated BRDF directions (W,S,3)/(W,D,3); ``brdf``:
    (n_brdf,S,D,B) already multiplied by pi; ``air``: (B,).  Returns a dict of
    every intermediate the golden fixtures hold.

    ``brdf_set_before_bake=False`` reproduces the reference when no BRDF / air
    attenuation was set before ``bake_geometry``: the baked tilde is then the bare
    form factor (RadiosityFast.py:415-422, :1260-1270) while ``init_source_energy``
    installs a unit BRDF stored as 1*pi (:459-473, :815) that only scales E0.
    """
    out = {}
    pts, ids = process_patches(walls_points, patch_size)
    cen, ar = centers(pts), areas(pts)
    nrm = _f64(walls_normal)[ids]
    out.update(patches_points=pts, patch_to_wall_ids=ids, patches_center=cen,
               patches_area=ar, patches_normal=nrm)
    vis = visibility_p2p(cen, nrm, pts)
    pairs = visible_pairs(vis)
    ff = ff_pairs(pts, nrm, ar, pairs)
    out.update(visibility=vis, visible_patches=pairs, ff_pairs=ff)
    if brdf_set_before_bake:
        tilde, odir, idir, delay = pair_tables(cen, ar, ids, pairs, ff, air, vi, vo,
                                               brdf, brdf_index, c, dt)
    else:
        tilde, odir, idir, delay = pair_tables(
            cen, ar, ids, pairs, ff, np.zeros(1), vi, vo, np.ones((1, 1, 1, 1)),
            np.zeros(len(_f64(walls_points)), np.int64), c, dt)
    out.update(tilde_pairs=tilde, out_dir=odir, in_dir=idir, pair_delays=delay)
    svis = visibility_pt2p(source, cen, walls_normal, walls_points)
    e0b, d0 = source_energy(source, cen, pts, svis, air)
    e0 = add_directional(e0b, source, cen, ids, vi, vo, brdf, brdf_index)
    out.update(source_visibility=svis, energy_0=e0b, distance_patches_to_source=d0,
               energy_init_source=e0)
    n_samples = int(duration / dt)
    etc = energy_exchange(e0, d0, pairs, tilde, odir, delay, n_samples, c, dt,
                          max_order, n_threads=n_threads)
    out["etc"] = etc
    monos, rvis, rfac, ridx, rdel = [], [], [], [], []
    for r in np.atleast_2d(receivers):
        v = visibility_pt2p(r, cen, walls_normal, walls_points)
        f = receiver_factor(r, pts, v)
        k = receiver_dir_index(cen, r, vo, ids)
        mono, _, dl = collect_receiver(etc, r, cen, f, k, air, c, dt)
        monos.append(mono), rvis.append(v), rfac.append(f), ridx.append(k)
        rdel.append(dl)
    out.update(etc_receiver_mono=np.array(monos), receiver_visibility=np.array(rvis),
               receiver_factor=np.array(rfac), receiver_dir_index=np.array(ridx),
               receiver_delays=np.array(rdel))
    return out
