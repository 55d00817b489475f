"""Host side of geometry baking: thin wrappers over the C ABI (csrc/bake.cu).

Each function mirrors one array-level operator of the reference (named in its
docstring) and takes/returns CUDA tensors.
"""
import ctypes

import torch

from . import _lib


def _dev(t, dtype=None):
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def make_blockers(surf_points, surf_normals):
    """Pre-digest blocking quadrilaterals ([M,4,3], [M,3]) for the visibility kernels."""
    lib = _lib.load()
    lib.spb_blocker_bytes.restype = ctypes.c_size_t
    m = surf_points.shape[0]
    if surf_points.shape[1] != 4:
        raise _lib.SparrowB200Error("only quadrilateral surfaces are supported")
    nbytes = lib.spb_blocker_bytes(ctypes.c_int64(m))
    buf = torch.empty(max(nbytes, 8) // 8, dtype=torch.float64, device=surf_points.device)
    _lib.call("spb_make_blockers", _dev(surf_points, torch.float64),
              _dev(surf_normals, torch.float64), m, _lib.I32(4), buf, _lib.stream_ptr())
    return buf


def visibility_p2p(centers, surf_normals, surf_points):
    """``geometry._check_patch2patch_visibility`` (reference geometry.py:750-797).

    Returns the (N, N) bool matrix, upper triangle only.
    """
    centers = _dev(centers, torch.float64)
    n, m = centers.shape[0], surf_points.shape[0]
    blockers = make_blockers(surf_points, surf_normals)
    vis = torch.empty((n, n), dtype=torch.uint8, device=centers.device)
    _lib.call("spb_visibility_p2p", centers, n, blockers, m, vis, _lib.stream_ptr())
    return vis.bool()


def visibility_pt2p(points, centers, surf_normals, surf_points, blockers=None):
    """``geometry._check_point2patch_visibility`` (geometry.py:799-839), batched over
    evaluation points: returns (R, N) bool."""
    points = _dev(points.reshape(-1, 3), torch.float64)
    centers = _dev(centers, torch.float64)
    n, m, r = centers.shape[0], surf_points.shape[0], points.shape[0]
    if blockers is None:
        blockers = make_blockers(surf_points, surf_normals)
    vis = torch.empty((r, n), dtype=torch.uint8, device=centers.device)
    _lib.call("spb_visibility_pt2p", points, r, centers, n, blockers, m, vis,
              _lib.stream_ptr())
    return vis.bool()


def visible_pairs(vis):
    """Row-major (P, 2) int32 list of visible pairs (RadiosityFast.py:377-387)."""
    return torch.nonzero(vis).to(torch.int32).contiguous()


def form_factors(points, normals, areas, pairs):
    """``patch2patch_ff_universal`` (universal.py:12-52) per visible pair: (P,) f64."""
    points, normals = _dev(points, torch.float64), _dev(normals, torch.float64)
    areas, pairs = _dev(areas, torch.float64), _dev(pairs, torch.int32)
    p = pairs.shape[0]
    ff = torch.zeros(p, dtype=torch.float64, device=points.device)
    flag = torch.zeros(p, dtype=torch.uint8, device=points.device)
    _lib.call("spb_form_factors_stokes", points, areas, pairs, p, ff, flag,
              _lib.stream_ptr())
    todo = torch.nonzero(flag).reshape(-1).contiguous()
    _lib.call("spb_form_factors_nusselt", points, normals, pairs, todo, todo.numel(), ff,
              _lib.stream_ptr())
    return ff, flag.bool()


def pair_geometry(centers, patch_to_wall, pairs, vi, vo):
    """Distance (numpy 1-D norm model) per pair and BRDF direction indices per
    directed pair (RadiosityFast.py:403-414, :538-543, :1386-1390).

    vi / vo: (W, S, 3) / (W, D, 3) or None for a diffuse 1x1 BRDF.
    """
    centers = _dev(centers, torch.float64)
    pairs = _dev(pairs, torch.int32)
    p = pairs.shape[0]
    dev = centers.device
    dist = torch.empty(p, dtype=torch.float64, device=dev)
    out_dir = torch.empty(2 * p, dtype=torch.int32, device=dev)
    in_dir = torch.empty(2 * p, dtype=torch.int32, device=dev)
    n_in = 1 if vi is None else vi.shape[1]
    n_out = 1 if vo is None else vo.shape[1]
    _lib.call("spb_pair_geometry", centers, _dev(patch_to_wall, torch.int64), pairs, p,
              None if vi is None else _dev(vi, torch.float64), n_in,
              None if vo is None else _dev(vo, torch.float64), n_out, dist, out_dir,
              in_dir, _lib.stream_ptr())
    return dist, out_dir, in_dir


def delay_bins(dist, speed_of_sound, dt):
    """``int(d / c / dt)`` (RadiosityFast.py:1067-1068, :1135-1136): int32."""
    dist = _dev(dist, torch.float64)
    out = torch.empty(dist.shape, dtype=torch.int32, device=dist.device)
    _lib.call("spb_delay_bins", dist, dist.numel(), float(speed_of_sound), float(dt), out,
              _lib.stream_ptr())
    return out


def source_energy(source, centers, points, vis, air, patch_to_wall, vi, brdf, brdf_index,
                  n_out):
    """``_source2patch_energy_universal`` + ``_add_directional``
    (universal.py:98-147, RadiosityFast.py:988-1034).

    Returns (distance (N,), e0 (N, D, B), energy (N, B))."""
    centers = _dev(centers, torch.float64)
    dev = centers.device
    n, nb = centers.shape[0], air.shape[0]
    dist = torch.empty(n, dtype=torch.float64, device=dev)
    e0 = torch.empty((n, n_out, nb), dtype=torch.float64, device=dev)
    energy = torch.empty((n, nb), dtype=torch.float64, device=dev)
    _lib.call("spb_source_energy", _dev(source.reshape(3), torch.float64), centers,
              _dev(points, torch.float64), _dev(vis, torch.uint8), _dev(air, torch.float64),
              _dev(patch_to_wall, torch.int64), _dev(vi, torch.float64), vi.shape[1],
              _dev(brdf, torch.float64), _dev(brdf_index, torch.int64), n_out, nb, n, dist,
              e0, energy, _lib.stream_ptr())
    return dist, e0, energy


def receiver_factors(receivers, centers, points, vis, air, patch_to_wall, vo,
                     speed_of_sound, dt, n_samples):
    """Receiver side of ``_collect_energy_patches`` (RadiosityFast.py:711-748) for a
    batch of receivers.  Returns dict(factor, rdir, delay, shift, scale)."""
    centers = _dev(centers, torch.float64)
    receivers = _dev(receivers.reshape(-1, 3), torch.float64)
    dev = centers.device
    r, n, nb = receivers.shape[0], centers.shape[0], air.shape[0]
    out = dict(
        factor=torch.empty((r, n), dtype=torch.float64, device=dev),
        rdir=torch.empty((r, n), dtype=torch.int32, device=dev),
        delay=torch.empty((r, n), dtype=torch.int32, device=dev),
        shift=torch.empty((r, n), dtype=torch.int32, device=dev),
        scale=torch.empty((r, n, nb), dtype=torch.float64, device=dev))
    _lib.call("spb_receiver_factors", receivers, r, centers, _dev(points, torch.float64),
              _dev(vis, torch.uint8), _dev(air, torch.float64),
              _dev(patch_to_wall, torch.int64), _dev(vo, torch.float64), vo.shape[1], nb, n,
              float(speed_of_sound), float(dt), int(n_samples), out["factor"], out["rdir"],
              out["delay"], out["shift"], out["scale"], _lib.stream_ptr())
    return out


def probe_norms(v):
    """x87-model Euclidean norm of each row of v ((n,2) or (n,3)) -- test probe."""
    v = _dev(v, torch.float64)
    out = torch.empty(v.shape[0], dtype=torch.float64, device=v.device)
    _lib.call("spb_probe_norms", v, v.shape[0], _lib.I32(v.shape[1]), out,
              _lib.stream_ptr())
    return out


def probe_basic_visibility(a, b, surf_points, surf_normals):
    """Element-wise ``_basic_visibility`` / ``_point_in_polygon`` -- test probe."""
    a, b = _dev(a, torch.float64), _dev(b, torch.float64)
    blockers = make_blockers(surf_points, surf_normals)
    n = a.shape[0]
    outs = [torch.empty(n, dtype=torch.uint8, device=a.device) for _ in range(3)]
    _lib.call("spb_probe_basic_visibility", a, b, blockers, n, outs[0], outs[1], outs[2],
              _lib.stream_ptr())
    return [o.bool() for o in outs]
