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


def visible_pairs(vis, max_block_elems=1 << 30):
    """Row-major (P, 2) int32 list of visible pairs (RadiosityFast.py:377-387).

    The matrix is scanned in blocks of rows: ``torch.nonzero`` handles at most 2^31
    elements per call, and config 5 (N = 100 000) has 10^10."""
    n_rows, n_cols = vis.shape
    step = max(1, int(max_block_elems // max(1, n_cols)))
    if step >= n_rows:
        return torch.nonzero(vis).to(torch.int32).contiguous()
    parts = []
    for r0 in range(0, n_rows, step):
        idx = torch.nonzero(vis[r0:r0 + step]).to(torch.int32)
        idx[:, 0] += r0
        parts.append(idx)
    return torch.cat(parts).contiguous()


def form_factors(points, normals, areas, pairs):
    """``patch2patch_ff_universal`` (universal.py:12-52) per visible pair: (P,) f64."""
    points, normals = _dev(points, torch.float64), _dev(normals, torch.float64)
    areas, pairs = _dev(areas, torch.float64), _dev(pairs, torch.int32)
    p = pairs.shape[0]
    ff = torch.zeros(p, dtype=torch.float64, device=points.device)
    flag = torch.zeros(p, dtype=torch.uint8, device=points.device)
    if p == 0:                       # e.g. a single plane: coplanar patches never see each other
        return ff, flag.bool()
    _lib.call("spb_form_factors_stokes", points, areas, pairs, p, ff, flag,
              _lib.stream_ptr())
    todo = torch.nonzero(flag).reshape(-1).contiguous()
    if todo.numel():
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
    if p == 0:
        return dist, out_dir, in_dir
    _lib.call("spb_pair_geometry", centers, _dev(patch_to_wall, torch.int64), pairs, p,
              None if vi is None else _dev(vi, torch.float64), n_in,
              None if vo is None else _dev(vo, torch.float64), n_out, dist, out_dir,
              in_dir, _lib.stream_ptr())
    return dist, out_dir, in_dir


def delay_bins(dist, speed_of_sound, dt):
    """``int(d / c / dt)`` (RadiosityFast.py:1067-1068, :1135-1136): int32."""
    dist = _dev(dist, torch.float64)
    out = torch.empty(dist.shape, dtype=torch.int32, device=dist.device)
    if dist.numel() == 0:
        return out
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


def source_energy_batch(sources, centers, points, vis, air, patch_to_wall, vi, brdf,
                        brdf_index, n_out):
    """:func:`source_energy` for S source positions in one launch: sources (S, 3), vis (S, N).
    Returns (distance (S, N), e0 (S, N, D, B))."""
    centers = _dev(centers, torch.float64)
    dev = centers.device
    sources = _dev(sources.reshape(-1, 3), torch.float64)
    n_src, n, nb = sources.shape[0], centers.shape[0], air.shape[0]
    dist = torch.empty((n_src, n), dtype=torch.float64, device=dev)
    e0 = torch.empty((n_src, n, n_out, nb), dtype=torch.float64, device=dev)
    for s0 in range(0, n_src, 65535):
        s1 = min(n_src, s0 + 65535)
        _lib.call("spb_source_energy_batch", sources[s0:s1], s1 - s0, centers,
                  _dev(points, torch.float64), _dev(vis[s0:s1], torch.uint8),
                  _dev(air, torch.float64), _dev(patch_to_wall, torch.int64),
                  _dev(vi, torch.float64), vi.shape[1], _dev(brdf, torch.float64),
                  _dev(brdf_index, torch.int64), n_out, nb, n, dist[s0:s1], e0[s0:s1], None,
                  _lib.stream_ptr())
    return dist, e0


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


# ---------------------------------------------------------------------------
# hierarchical visibility (csrc/vis_group.cuh)
# ---------------------------------------------------------------------------
_BLK = dict(n=slice(0, 3), s0=slice(3, 6), r0=slice(6, 9), r1=slice(9, 12), xmin=32, xmax=33,
            ymin=34, ymax=35, h=36, aa2d=37)
# exact::Group (csrc/vis_group.cuh): 17 doubles, then int32 fields from byte 136
_GRP_BYTES = 176
_GRP_D = dict(n=slice(0, 3), s0=slice(3, 6), r0=slice(6, 9), r1=slice(9, 12), plane_dev=12,
              y0=13, inv_bin_h=14, x0=15, inv_bin_w=16)
_GRP_I = dict(n_bins=34, bin_ptr0=35, m0=36, m1=37, n_bx=38, cell_ptr0=39, strip0=40,
              n_strips=41, sfirst0=42)
_BIN_MARGIN = 2e-3          # members are listed with this margin around their ray band
_MIN_CELL_MEMBERS = 16      # smaller groups keep the 1-D bins only


def _csr_lists(first, last, n_lists, values):
    """CSR of `n_lists` lists where element k (payload values[k]) is a member of the lists
    first[k] .. last[k]; within a list the elements keep their order.  Returns (ptr, items)."""
    import numpy as np
    cnt = np.maximum(last - first + 1, 0)
    rep = np.repeat(np.arange(len(cnt)), cnt)
    off = np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    lst = first[rep] + off
    order = np.argsort(lst, kind="stable")
    ptr = np.concatenate([[0], np.cumsum(np.bincount(lst, minlength=n_lists))])
    return ptr.astype(np.int64), values[rep][order]


def build_groups(blockers_np, group_ids):
    """Group table for ``spb_visibility_p2p_grouped`` (host-side index bookkeeping).

    blockers_np: (M, 38) float64 view of the Blocker records; group_ids: (M,) the wall
    each blocker belongs to.  A wall becomes one group when all its blockers share the
    bitwise-same normal (hence rotation) and their first vertices lie in one plane to
    within 1e-12; otherwise its blockers become singleton groups (always correct, just
    not accelerated).  Per group: bins along the in-plane y axis listing every member of the
    band, and -- for groups of >= 16 members -- a 2-D grid of cells listing the members whose
    expanded bounding box touches the cell, plus the y-ranges ("strips") in which a point far
    to the left of an axis-aligned member can still be a candidate (the kernel then scans the
    bin instead of the cell).  Returns numpy arrays (groups uint8 (G, group_bytes), members
    int32, bin_ptr int32, bin_items int32, strips float64 (n, 2), group_of int32 (M,) = the
    group of every blocker).
    """
    import numpy as np
    lib = _lib.load()
    lib.spb_group_bytes.restype = ctypes.c_size_t
    gbytes = int(lib.spb_group_bytes())
    assert gbytes == _GRP_BYTES, gbytes
    blk = np.ascontiguousarray(blockers_np, dtype=np.float64)
    ids = np.asarray(group_ids)
    member_lists = []
    order = np.argsort(ids, kind="stable")
    cuts = np.nonzero(np.diff(ids[order]))[0] + 1
    for idx in np.split(order, cuts):
        key = blk[idx][:, :12].copy()
        key[:, 3:6] = 0.0                                 # n, r0, r1 must be bitwise equal
        same = (key.view(np.uint64) == key[:1].view(np.uint64)).all()
        n0, s00 = blk[idx[0], _BLK["n"]], blk[idx[0], _BLK["s0"]]
        dev = np.abs((blk[idx][:, _BLK["s0"]] - s00) @ n0).max()
        if same and dev <= 1e-12:
            member_lists.append((idx, float(dev) + 1e-13))
        else:
            member_lists.extend((idx[k:k + 1], 0.0) for k in range(len(idx)))
    n_groups = len(member_lists)
    groups = np.zeros((n_groups, gbytes), dtype=np.uint8)
    gd = groups.view(np.float64).reshape(n_groups, gbytes // 8)
    gi = groups.view(np.int32).reshape(n_groups, gbytes // 4)
    members, ptr_parts, item_parts, strip_parts = [], [], [], []
    n_ptr = n_items = n_members = n_strips_all = 0

    def add_lists(ptr, items):
        """Append one CSR block to bin_ptr / bin_items; returns its offset into bin_ptr."""
        nonlocal n_ptr, n_items
        at = n_ptr
        ptr_parts.append(ptr + n_items)                   # len(lists) + 1 entries
        item_parts.append(items)
        n_ptr += len(ptr)
        n_items += len(items)
        return at

    for g, (idx, dev) in enumerate(member_lists):
        b = blk[idx]
        idx32 = idx.astype(np.int32)
        h = b[:, _BLK["h"]]
        lo = b[:, _BLK["ymin"]] - h - _BIN_MARGIN
        hi = b[:, _BLK["ymax"]] + h + _BIN_MARGIN
        bin_h = float((hi - lo).max())
        y0 = float(lo.min())
        n_bins = int(np.floor((hi.max() - y0) / bin_h)) + 1
        first = np.floor((lo - y0) / bin_h).astype(np.int64)
        last = np.minimum(np.floor((hi - y0) / bin_h).astype(np.int64), n_bins - 1)
        # one extra bin on each side guards the rounding of the kernel's bin index
        first, last = np.maximum(first - 1, 0), np.minimum(last + 1, n_bins - 1)
        ptr, items = _csr_lists(first, last, n_bins, idx32)
        gd[g, _GRP_D["n"]], gd[g, _GRP_D["s0"]] = b[0, _BLK["n"]], b[0, _BLK["s0"]]
        gd[g, _GRP_D["r0"]], gd[g, _GRP_D["r1"]] = b[0, _BLK["r0"]], b[0, _BLK["r1"]]
        gd[g, _GRP_D["plane_dev"]], gd[g, _GRP_D["y0"]] = dev, y0
        gd[g, _GRP_D["inv_bin_h"]] = 1.0 / bin_h
        gi[g, _GRP_I["n_bins"]], gi[g, _GRP_I["bin_ptr0"]] = n_bins, add_lists(ptr, items)
        gi[g, _GRP_I["m0"]], gi[g, _GRP_I["m1"]] = n_members, n_members + len(idx)
        members.append(idx32)
        n_members += len(idx)
        if len(idx) < _MIN_CELL_MEMBERS:
            continue
        # --- 2-D cells: members whose expanded box [xlo, xhi] x [lo, hi] touches the cell.  A
        # member that is not an exactly axis-aligned rectangle is a candidate for every point
        # to its left (ray rule only), so it is listed in all the cells left of it as well.
        aa = b[:, _BLK["aa2d"]] != 0.0
        xlo = b[:, _BLK["xmin"]] - _BIN_MARGIN
        xhi = b[:, _BLK["xmax"]] + h + _BIN_MARGIN
        bin_w = float((xhi - xlo).max())
        x0 = float(xlo.min())
        inv_w, inv_h = 1.0 / bin_w, gd[g, _GRP_D["inv_bin_h"]]
        n_bx = int(np.floor((xhi.max() - x0) * inv_w)) + 1
        if n_bx * n_bins > 4 * len(idx) + 64:              # degenerate spread: not worth it
            continue
        # the kernel computes floor((q - origin) * inv) with the same two IEEE operations, which
        # are monotonic in q; the extra 1e-6 makes ulp-level differences irrelevant
        bx0 = np.where(aa, np.floor((xlo - 1e-6 - x0) * inv_w), 0.0).astype(np.int64)
        bx1 = np.floor((xhi + 1e-6 - x0) * inv_w).astype(np.int64)
        by0 = np.floor((lo - 1e-6 - y0) * inv_h).astype(np.int64)
        by1 = np.floor((hi + 1e-6 - y0) * inv_h).astype(np.int64)
        bx0, bx1 = np.clip(bx0, 0, n_bx - 1), np.clip(bx1, 0, n_bx - 1)
        by0, by1 = np.clip(by0, 0, n_bins - 1), np.clip(by1, 0, n_bins - 1)
        # expand over the rows first, then over the columns of each (member, row)
        rptr, mem_row = _csr_lists(by0, by1, n_bins, np.arange(len(idx)))
        rows_of = np.repeat(np.arange(n_bins), np.diff(rptr))
        cell_first = rows_of * n_bx + bx0[mem_row]
        cell_last = rows_of * n_bx + bx1[mem_row]
        cptr, citems = _csr_lists(cell_first, cell_last, n_bx * n_bins, idx32[mem_row])
        # --- strips: y-ranges where a point left of an axis-aligned member grazes an edge
        sl = np.concatenate([lo[aa], b[aa, _BLK["ymax"]] - _BIN_MARGIN])
        sh = np.concatenate([b[aa, _BLK["ymin"]] + _BIN_MARGIN, hi[aa]])
        o = np.argsort(sl, kind="stable")
        sl, sh = sl[o], sh[o]
        if len(sl):
            run_hi = np.maximum.accumulate(sh)
            start = np.concatenate([[True], sl[1:] > run_hi[:-1]])      # a gap before strip k
            seg = np.cumsum(start) - 1
            s_lo = sl[start]
            s_hi = np.full(len(s_lo), -np.inf)
            np.maximum.at(s_hi, seg, sh)
        else:
            s_lo = s_hi = np.zeros(0)
        edges = y0 + (np.arange(n_bins) * bin_h) - 1e-6    # lower edge of each bin, with slack
        sfirst = np.searchsorted(s_hi, edges, side="left").astype(np.int64)
        gd[g, _GRP_D["x0"]], gd[g, _GRP_D["inv_bin_w"]] = x0, inv_w
        gi[g, _GRP_I["n_bx"]], gi[g, _GRP_I["cell_ptr0"]] = n_bx, add_lists(cptr, citems)
        gi[g, _GRP_I["strip0"]], gi[g, _GRP_I["n_strips"]] = n_strips_all, len(s_lo)
        # the per-bin first-strip indices ride in bin_ptr as a list block without items
        ptr_parts.append(sfirst)
        gi[g, _GRP_I["sfirst0"]] = n_ptr
        n_ptr += len(sfirst)
        strip_parts.append(np.stack([s_lo, s_hi], 1))
        n_strips_all += len(s_lo)
    assert n_ptr < 2 ** 31 and n_items < 2 ** 31
    cat = lambda parts, dt: (np.concatenate(parts).astype(dt) if parts  # noqa: E731
                             else np.zeros(0, dt))
    bin_items = cat(item_parts, np.int32)
    strips = np.concatenate(strip_parts + [np.zeros((1, 2))])           # never empty
    group_of = np.full(len(ids), -1, np.int32)
    for g, (idx, _) in enumerate(member_lists):
        group_of[idx] = g
    return (groups, cat(members, np.int32), cat(ptr_parts, np.int32),
            bin_items if len(bin_items) else np.zeros(1, np.int32),
            np.ascontiguousarray(strips, dtype=np.float64), group_of)


def make_blockers_host(surf_points, surf_normals):
    """CPU twin of make_blockers (numpy in, (M, 38) float64 out) -- tests only."""
    import numpy as np
    lib = _lib.load()
    pts = np.ascontiguousarray(surf_points, dtype=np.float64)
    nrm = np.ascontiguousarray(surf_normals, dtype=np.float64)
    lib.spb_blocker_bytes.restype = ctypes.c_size_t
    m = pts.shape[0]
    out = np.zeros((m, lib.spb_blocker_bytes(ctypes.c_int64(m)) // 8 // max(m, 1)))
    rc = lib.spb_make_blockers_host(pts.ctypes.data_as(ctypes.c_void_p),
                                    nrm.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(m),
                                    out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def _own_group(group_of, n):
    """Per centre i the group of blocker i (-1 where there is no blocker i)."""
    import numpy as np
    out = np.full(n, -1, np.int32)
    k = min(n, len(group_of))
    out[:k] = group_of[:k]
    return out


def visibility_p2p_grouped_host(centers, surf_normals, surf_points, group_ids, hints=True):
    """CPU twin of the grouped visibility kernel (numpy) -- tests only.  ``hints=False``
    switches the memoised own-polygon test and the own-walls-first order off."""
    import numpy as np
    lib = _lib.load()
    cen = np.ascontiguousarray(centers, dtype=np.float64)
    blk = make_blockers_host(surf_points, surf_normals)
    groups, members, bin_ptr, bin_items, strips, group_of = build_groups(blk, group_ids)
    n = cen.shape[0]
    own_group = _own_group(group_of, n)
    vis = np.zeros((n, n), np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    rc = lib.spb_visibility_p2p_grouped_host(
        p(cen), ctypes.c_int64(n), p(blk), p(groups), ctypes.c_int64(len(groups)), p(members),
        p(bin_ptr), p(bin_items), p(strips), ctypes.c_int64(min(n, len(blk)) if hints else 0),
        p(own_group) if hints else None, p(vis))
    assert rc == 0
    return vis.astype(bool)


def visibility_p2p_grouped(centers, surf_normals, surf_points, group_ids, row_range=None):
    """``geometry._check_patch2patch_visibility`` (geometry.py:750-797), hierarchical:
    same (N, N) bool matrix as :func:`visibility_p2p` in O(N^2 * walls).  ``row_range`` =
    ``(lo, hi)`` computes those rows only and returns a (hi - lo, N) matrix -- the unit of a
    bake that is sharded over GPUs."""
    centers = _dev(centers, torch.float64)
    n = centers.shape[0]
    blockers = make_blockers(surf_points, surf_normals)
    m = surf_points.shape[0]
    blk_np = blockers.cpu().numpy().reshape(m, -1)
    groups, members, bin_ptr, bin_items, strips, group_of = build_groups(
        blk_np, group_ids.cpu().numpy() if isinstance(group_ids, torch.Tensor) else group_ids)
    dev = centers.device
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    # memo "centre i lies in polygon i" and the group of blocker i (result-neutral hints)
    own_in = torch.full((n,), 255, dtype=torch.uint8, device=dev)
    _lib.call("spb_visibility_own_in", centers, min(n, m), blockers, own_in, _lib.stream_ptr())
    own_group = t(_own_group(group_of, n))
    if row_range is None:
        vis = torch.empty((n, n), dtype=torch.uint8, device=dev)
        _lib.call("spb_visibility_p2p_grouped", centers, n, blockers, t(groups), len(groups),
                  t(members), t(bin_ptr), t(bin_items), t(strips), own_in, own_group, vis,
                  _lib.stream_ptr())
        return vis.bool()
    lo, hi = int(row_range[0]), int(row_range[1])
    vis = torch.empty((hi - lo, n), dtype=torch.uint8, device=dev)
    _lib.call("spb_visibility_p2p_grouped_rows", centers, n, blockers, t(groups), len(groups),
              t(members), t(bin_ptr), t(bin_items), t(strips), own_in, own_group, lo, hi, vis,
              _lib.stream_ptr())
    return vis.bool()


def triangle_row_range(n, part, n_parts):
    """Rows of the upper-triangular visibility matrix for part ``part`` of ``n_parts`` with
    (nearly) equal numbers of entries: row i holds n - 1 - i pairs."""
    import numpy as np
    total = n * (n - 1) // 2
    rows = np.arange(n + 1, dtype=np.int64)
    before = rows * (2 * n - rows - 1) // 2          # pairs in rows [0, i)
    lo = int(np.searchsorted(before, total * part // n_parts, side="left"))
    hi = int(np.searchsorted(before, total * (part + 1) // n_parts, side="left"))
    if part == n_parts - 1:
        hi = n
    return min(lo, n), min(hi, n)
