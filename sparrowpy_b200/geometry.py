"""Wall polygons and patch tessellation (host side).

``Polygon`` mirrors the reference's input type (reference geometry.py:15-230) and
the functions below restate its patch bookkeeping (geometry.py:290-448) in numpy
float64 -- plain IEEE operations in the same order, so the results are identical.
This stage costs milliseconds; it feeds the CUDA baking kernels.
"""
import numpy as np


class Polygon:
    """Planar convex polygon given by its corner points, up vector and normal."""

    def __init__(self, points, up_vector, normal):
        self.pts = np.array(points, dtype=float)
        normal = np.array(normal, dtype=float)
        assert self.pts.shape[1] >= 3, \
            'You need at least 3 points to build a Polygon'
        if self.n_points > 3:
            x_0 = self.pts[0]
            for i in range(1, self.n_points - 2):
                det = np.linalg.det(
                    [x_0 - self.pts[i], x_0 - self.pts[i + 1], x_0 - self.pts[i + 2]])
                assert abs(det) < 1e-12, \
                    'Points must be in a plane to create a Polygon'
        up = np.array(up_vector, dtype=float)
        self.up_vector = up / np.sqrt(np.dot(up, up))
        calc = np.cross(self.pts[0] - self.pts[1], self.pts[0] - self.pts[2])
        calc = calc / np.sqrt(np.dot(calc, calc))
        assert all(np.abs(np.cross(normal, calc)) < 1e-10), \
            'The normal vector is not perpendicular to the polygon'
        self._normal = normal

    @property
    def normal(self):
        return self._normal

    @property
    def n_points(self):
        return self.pts.shape[0]

    @property
    def center(self):
        return np.sum(self.pts, axis=0) / self.n_points

    @property
    def size(self):
        vec1 = self.pts[0] - self.pts[1]
        vec2 = self.pts[1] - self.pts[2]
        return np.abs(vec1 - vec2)

    @property
    def area(self):
        return float(calculate_area(self.pts[None])[0])

    def to_dict(self):
        return {'up_vector': self.up_vector.tolist(), 'pts': self.pts.tolist(),
                'normal': self._normal.tolist()}

    @classmethod
    def from_dict(cls, input_dict):
        return cls(input_dict['pts'], input_dict['up_vector'], input_dict['normal'])


def _wall_grid(wall, max_size):
    """Patch counts and in-plane axes of one wall (geometry.py:357-369)."""
    size = wall.max(axis=0) - wall.min(axis=0)
    with np.errstate(divide='ignore', invalid='ignore'):
        nums = np.array([int(n) for n in size / max_size])
    x_idx, y_idx = 0, 1
    if nums[2] == 0:
        x_idx, y_idx = 0, 1
    if nums[1] == 0:
        x_idx, y_idx = 0, 2
    if nums[0] == 0:
        x_idx, y_idx = 1, 2
    return size, nums, x_idx, y_idx


def process_patches(walls_points, patch_size):
    """Tessellate axis-aligned rectangular walls (geometry.py:290-410).

    Returns ``(patches_points (N,4,3), patch_to_wall_ids (N,) int64)``.
    """
    walls_points = np.asarray(walls_points, dtype=float)
    patches, ids = [], []
    for w, wall in enumerate(walls_points):
        size, nums, xi, yi = _wall_grid(wall, patch_size)
        nx, ny = int(nums[xi]), int(nums[yi])
        if nx * ny == 0:
            continue
        rsx, rsy = size[xi] / nx, size[yi] / ny
        x_min, y_min = wall[:, xi].min(), wall[:, yi].min()
        ix = np.repeat(np.arange(nx), ny).astype(float)      # i_x outer, i_y inner
        iy = np.tile(np.arange(ny), nx).astype(float)
        pts = np.broadcast_to(wall, (nx * ny, 4, 3)).copy()
        x0, x1 = x_min + ix * rsx, x_min + (ix + 1) * rsx
        y0, y1 = y_min + iy * rsy, y_min + (iy + 1) * rsy
        pts[:, 0, xi], pts[:, 0, yi] = x0, y0
        pts[:, 1, xi], pts[:, 1, yi] = x1, y0
        pts[:, 3, xi], pts[:, 3, yi] = x0, y1
        pts[:, 2, xi], pts[:, 2, yi] = x1, y1
        patches.append(pts)
        ids.append(np.full(nx * ny, w, dtype=np.int64))
    if not patches:
        return np.empty((0, 4, 3)), np.empty(0, np.int64)
    return np.concatenate(patches), np.concatenate(ids)


def calculate_center(points):
    """geometry.py:412-413 (sequential sum over the vertex axis, then /n)."""
    points = np.asarray(points, dtype=float)
    s = np.zeros(points.shape[:-2] + (3,))
    for k in range(points.shape[-2]):
        s = s + points[..., k, :]
    return s / points.shape[-2]


def calculate_size(points):
    """geometry.py:415-418"""
    vec1 = points[..., 0, :] - points[..., 1, :]
    vec2 = points[..., 1, :] - points[..., 2, :]
    return np.abs(vec1 - vec2)


def calculate_area(points):
    """geometry.py:420-448: triangle fan, 0.5*|cross|.

    The reference's norm is the x87 80-bit dnrm2; numpy's longdouble
    is that format on x86-64, elsewhere this is a <=1 ulp tolerance path.
    """
    points = np.asarray(points, dtype=float)
    area = np.zeros(points.shape[0])
    for t in range(points.shape[1] - 2):
        c = np.cross(points[:, t + 1] - points[:, 0], points[:, t + 2] - points[:, 0])
        cl = c.astype(np.longdouble)
        nrm = np.sqrt((cl[:, 0] * cl[:, 0] + cl[:, 1] * cl[:, 1]) + cl[:, 2] * cl[:, 2])
        area = area + .5 * nrm.astype(np.float64)
    return area


def compact_patch_order(patches_points, patch_to_wall_ids, block=(2, 4), n_shards=1):
    """Internal patch numbering in which every run of 8 consecutive indices is a
    compact ``block`` (2 x 4 patches) of one wall instead of an 8 x 1 strip.

    Used only to lay out the receiver tiles of the energy-exchange kernel (any
    numbering gives the same result).  Blocks at the rim of a wall grid are not full,
    so the internal index space has holes: returns ``(rank, n_internal)`` with
    ``rank[patch] = internal index`` and ``n_internal = 8 * number of blocks >= N``.

    ``n_shards`` > 1 deals the blocks round-robin to that many equal contiguous index
    ranges (the receiver shards of a multi-GPU run), so that every shard gets the
    same mix of walls and therefore a similar number of visible pairs.
    """
    pts = np.asarray(patches_points, dtype=float)
    ids = np.asarray(patch_to_wall_ids)
    center = pts.mean(axis=1)
    per = block[0] * block[1]
    rank = np.empty(len(ids), dtype=np.int64)
    base = 0
    for w in np.unique(ids):
        sel = np.nonzero(ids == w)[0]
        ext = pts[sel].max(axis=(0, 1)) - pts[sel].min(axis=(0, 1))
        axes = sorted(np.argsort(ext)[1:])               # the two in-plane axes
        size = (pts[sel].max(axis=1) - pts[sel].min(axis=1)).max(axis=0)
        idx = []
        for ax in axes:
            step = size[ax] if size[ax] > 0 else 1.0
            idx.append(np.rint((center[sel, ax] - center[sel, ax].min()) / step).astype(
                np.int64))
        nb1 = int(idx[1].max()) // block[1] + 1
        blk = (idx[0] // block[0]) * nb1 + idx[1] // block[1]
        within = (idx[0] % block[0]) * block[1] + idx[1] % block[1]
        rank[sel] = base + blk * per + within
        base += (int(blk.max()) + 1) * per
    n_internal = int(base)
    if n_shards > 1:
        n_tiles = n_internal // per
        per_shard = -(-n_tiles // n_shards)
        tile = rank // per
        rank = ((tile % n_shards) * per_shard + tile // n_shards) * per + rank % per
        n_internal = n_shards * per_shard * per
    assert len(np.unique(rank)) == len(rank)
    return rank, n_internal
