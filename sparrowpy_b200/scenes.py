"""Synthetic polygon scenes for the configurations named in BASELINE.json.

Every wall is an axis-aligned rectangle with integer-metre extents, because the
reference tessellation only supports those (reference geometry.py:341-410, the
zero-extent axis is found from ``int(size/max_size) == 0``).

A wall is a tuple ``(points(4,3), up_vector(3,), normal(3,))`` -- the argument
order of ``Polygon(points, up_vector, normal)`` (reference geometry.py:25-27).
"""
import numpy as np

SPEED_OF_SOUND = 343.2
ETC_DT = 1e-3


def _rect(axis, const, lo, hi, normal_sign):
    """Axis-aligned rectangle lying in the plane ``axis == const``.

    ``lo``/``hi`` are the 2-vectors of the two remaining axes in (x, y, z) order.
    Vertex order: (lo0, lo1), (hi0, lo1), (hi0, hi1), (lo0, hi1).
    """
    others = [a for a in range(3) if a != axis]
    corners = [(lo[0], lo[1]), (hi[0], lo[1]), (hi[0], hi[1]), (lo[0], hi[1])]
    pts = np.zeros((4, 3))
    for k, (u, v) in enumerate(corners):
        pts[k, axis] = const
        pts[k, others[0]] = u
        pts[k, others[1]] = v
    normal = np.zeros(3)
    normal[axis] = normal_sign
    up = np.zeros(3)
    # the up vector lies in the wall plane: z for vertical walls, x for floors
    up[2 if axis != 2 else 0] = 1.0
    return pts, up, normal


def shoebox(lx, ly, lz):
    """Six inward-facing walls of a shoebox room.

    Wall and vertex order follow the reference's test stub
    (reference testing/stub_utils.py:5-48) so that patch numbering is identical.
    """
    lx, ly, lz = float(lx), float(ly), float(lz)
    walls = []
    # y = 0 and y = ly walls: vertices run (x, z)
    for y, ny in ((0.0, 1.0), (ly, -1.0)):
        pts = np.array([[0, y, 0], [lx, y, 0], [lx, y, lz], [0, y, lz]], float)
        walls.append((pts, np.array([1.0, 0, 0]), np.array([0, ny, 0.0])))
    # floor and ceiling: vertices run (x, y)
    for z, nz in ((0.0, 1.0), (lz, -1.0)):
        pts = np.array([[0, 0, z], [lx, 0, z], [lx, ly, z], [0, ly, z]], float)
        walls.append((pts, np.array([1.0, 0, 0]), np.array([0, 0, nz])))
    # x = 0 and x = lx walls: vertices run (z, y)
    for x, nx in ((0.0, 1.0), (lx, -1.0)):
        pts = np.array([[x, 0, 0], [x, 0, lz], [x, ly, lz], [x, ly, 0]], float)
        walls.append((pts, np.array([0, 0, 1.0]), np.array([nx, 0, 0.0])))
    return walls


def ground_plane(x0, x1, y0, y1, z=0.0):
    """Single upward-facing ground polygon (config 3)."""
    return [_rect(2, float(z), (float(x0), float(y0)), (float(x1), float(y1)), 1.0)]


def building(x0, y0, sx, sy, height, roof=False):
    """Outward-facing facades of a box building standing on z = 0."""
    x1, y1 = x0 + sx, y0 + sy
    walls = [
        _rect(0, float(x0), (float(y0), 0.0), (float(y1), float(height)), -1.0),
        _rect(0, float(x1), (float(y0), 0.0), (float(y1), float(height)), 1.0),
        _rect(1, float(y0), (float(x0), 0.0), (float(x1), float(height)), -1.0),
        _rect(1, float(y1), (float(x0), 0.0), (float(x1), float(height)), 1.0),
    ]
    if roof:
        walls.append(_rect(2, float(height), (float(x0), float(y0)),
                           (float(x1), float(y1)), 1.0))
    return walls


def occluder_scene(ground=6, box=2, height=2):
    """Small non-convex scene: ground square with a roofed box in the middle."""
    g0 = (ground - box) // 2
    return ground_plane(0, ground, 0, ground) + building(
        g0, g0, box, box, height, roof=True)


def _rows_of_buildings(n_rows_counts, gx, gy, b, seed):
    """Seeded integer building positions on a south and a north row.

    The rows leave a street along x in the middle of the ground.  Each row is cut
    into equal slots; a building sits in its slot with a seeded integer jitter.
    Returns a list of (x0, y0, width).
    """
    rng = np.random.default_rng(seed)
    edge = int(gy / 12)
    rows_y = [edge, gy - edge - b]
    placed = []
    for count, y0 in zip(n_rows_counts, rows_y):
        slot = gx // count
        width = min(b, slot - 1)
        for k in range(count):
            jitter = int(rng.integers(0, slot - width))
            placed.append((k * slot + jitter, y0, width))
    return placed


def street_canyon(seed=0, scale=1.0):
    """Config 4: ground 120x60 m + 5 buildings 20x20x30 m (20 facades, no roofs)
    on both sides of a street.

    ``scale`` < 1 shrinks every extent (kept integer) for parity-test sizes.
    """
    gx, gy = int(round(120 * scale)), int(round(60 * scale))
    b, h = max(1, int(round(20 * scale))), max(1, int(round(30 * scale)))
    walls = ground_plane(0, gx, 0, gy)
    for x0, y0, w in _rows_of_buildings((3, 2), gx, gy, b, seed):
        walls += building(x0, y0, w, b, h)
    return walls


def city_block(seed=0, scale=1.0):
    """Config 5: ground 120x75 m + 10 buildings 20x20x20 m (40 facades)."""
    gx, gy = int(round(120 * scale)), int(round(75 * scale))
    b = max(1, int(round(20 * scale)))
    walls = ground_plane(0, gx, 0, gy)
    for x0, y0, w in _rows_of_buildings((5, 5), gx, gy, b, seed):
        walls += building(x0, y0, w, b, b)
    return walls


def hemisphere_directions(n_azimuth=8, colatitudes_deg=(30.0, 60.0)):
    """Direction set of config 2: azimuths x colatitudes on the upper hemisphere.

    Returns unit vectors (n, 3) in the BRDF frame (normal +z, up +x) and weights
    normalised to 2*pi.
    """
    dirs = []
    for col in colatitudes_deg:
        th = np.deg2rad(col)
        for k in range(n_azimuth):
            ph = 2 * np.pi * k / n_azimuth
            dirs.append([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph),
                         np.cos(th)])
    dirs = np.array(dirs)
    weights = np.full(len(dirs), 2 * np.pi / len(dirs))
    return dirs, weights


def brdf_from_scattering(directions, weights, scattering, absorption):
    """Discretised BRDF of a surface with random-incidence scattering ``s`` and
    absorption ``alpha`` per band: specular lobe into the mirrored direction bin
    plus a Lambertian part (the formula documented at reference brdf.py:19-24).

    Returns (S, D, B) with S == D == len(directions); NOT yet multiplied by pi
    (``set_wall_brdf`` does that, reference RadiosityFast.py:815).
    """
    directions = np.asarray(directions, float)
    s = np.atleast_1d(np.asarray(scattering, float))
    a = np.atleast_1d(np.asarray(absorption, float))
    n = len(directions)
    w = np.asarray(weights, float) * (2 * np.pi / np.sum(weights))
    brdf = np.zeros((n, n, len(s)))
    brdf += s / np.pi
    mirrored = directions * np.array([-1.0, -1.0, 1.0])
    cos_in = directions[:, 2]
    for i in range(n):
        o = int(np.argmin(np.sum((directions - mirrored[i]) ** 2, axis=-1)))
        brdf[i, o, :] += (1 - s) / (cos_in[i] * w[o])
    brdf *= (1 - a)
    return brdf
