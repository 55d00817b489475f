"""Minimal stand-ins for the pyfar value types on the DirectionalRadiosityFast path.

pyfar is not installed in this image.  The reference API takes and returns
``pf.Coordinates`` / ``pf.FrequencyData`` / ``pf.TimeData`` (SURVEY.md section 8b
lists the attribute surface that is touched: ``.cartesian .cshape .csize .cdim .z
.weights .copy() .radius``; ``.freq .frequencies .cshape .n_bins``; ``.time
.times``).  When pyfar *is* importable the public class accepts the real types as
well (duck typing); these shims are what tests and bench use.
"""
import numpy as np


class Coordinates:
    """Cartesian point set with optional weights, cshape ``(n,)``."""

    def __init__(self, x=None, y=None, z=None, weights=None):
        if x is None:
            pts = np.zeros((0, 3))
        else:
            x, y, z = np.broadcast_arrays(np.atleast_1d(np.asarray(x, float)),
                                          np.atleast_1d(np.asarray(y, float)),
                                          np.atleast_1d(np.asarray(z, float)))
            pts = np.stack([x, y, z], axis=-1)
        self._pts = np.array(pts, dtype=float)
        self.weights = None if weights is None else np.broadcast_to(
            np.atleast_1d(np.asarray(weights, float)),
            self._pts.shape[:-1]).copy()

    @classmethod
    def from_cartesian(cls, xyz, weights=None):
        xyz = np.asarray(xyz, float)
        return cls(xyz[..., 0], xyz[..., 1], xyz[..., 2], weights=weights)

    @property
    def cartesian(self):
        return self._pts

    @cartesian.setter
    def cartesian(self, value):
        self._pts = np.array(value, dtype=float)

    @property
    def cshape(self):
        return self._pts.shape[:-1]

    @property
    def csize(self):
        return int(np.prod(self.cshape))

    @property
    def cdim(self):
        return len(self.cshape)

    @property
    def x(self):
        return self._pts[..., 0]

    @property
    def y(self):
        return self._pts[..., 1]

    @property
    def z(self):
        return self._pts[..., 2]

    @property
    def radius(self):
        return np.sqrt(np.sum(self._pts ** 2, axis=-1))

    @radius.setter
    def radius(self, value):
        r = self.radius
        self._pts = self._pts / r[..., None] * value

    @property
    def colatitude(self):
        return np.arccos(self._pts[..., 2] / self.radius)

    @property
    def azimuth(self):
        return np.mod(np.arctan2(self._pts[..., 1], self._pts[..., 0]),
                      2 * np.pi)

    @azimuth.setter
    def azimuth(self, value):
        r, col = self.radius, self.colatitude
        az = np.broadcast_to(np.asarray(value, float), r.shape)
        self._pts = np.stack([r * np.sin(col) * np.cos(az), r * np.sin(col) * np.sin(az),
                              r * np.cos(col)], axis=-1)

    def find_nearest(self, find):
        """Index of the nearest point of ``self`` for every point of ``find``:
        returns ``((index_array,), distance)`` like pyfar."""
        a = self._pts.reshape(-1, 3)
        b = find.cartesian.reshape(-1, 3)
        d2 = np.sum((b[:, None, :] - a[None, :, :]) ** 2, axis=-1)
        idx = np.argmin(d2, axis=1)
        return (idx,), np.sqrt(d2[np.arange(len(b)), idx])

    def copy(self):
        out = Coordinates.from_cartesian(self._pts.copy())
        out.weights = None if self.weights is None else self.weights.copy()
        return out

    def apply_matrix(self, matrix):
        """Rotate all points by a 3x3 matrix (column-vector convention)."""
        self._pts = self._pts @ np.asarray(matrix, float).T

    def __sub__(self, other):
        return Coordinates.from_cartesian(self._pts - other.cartesian)

    def __getitem__(self, idx):
        out = Coordinates.from_cartesian(self._pts[idx])
        if self.weights is not None:
            out.weights = self.weights[idx]
        return out


class FrequencyData:
    """Frequency-domain data: ``freq`` of shape ``(*cshape, n_bins)``."""

    def __init__(self, data, frequencies):
        self.frequencies = np.atleast_1d(np.asarray(frequencies, float))
        data = np.asarray(data)
        # pyfar stores complex spectra; every quantity on this path is real
        # (the reference takes np.real() of it, RadiosityFast.py:1032), so real
        # input stays real.
        self.freq = np.atleast_2d(data).astype(
            complex if np.iscomplexobj(data) else float)
        if self.freq.shape[-1] != self.frequencies.size:
            raise ValueError("Number of frequency values does not match the "
                             "number of frequencies")

    @property
    def n_bins(self):
        return self.freq.shape[-1]

    @property
    def cshape(self):
        return self.freq.shape[:-1]


class TimeData:
    """Time-domain data: ``time`` of shape ``(*cshape, n_samples)``."""

    def __init__(self, data, times):
        self.time = np.atleast_2d(np.asarray(data, float))
        self.times = np.atleast_1d(np.asarray(times, float))
        if self.time.shape[-1] != self.times.size:
            raise ValueError("The length of times must be data.shape[-1]")

    @property
    def n_samples(self):
        return self.time.shape[-1]

    @property
    def cshape(self):
        return self.time.shape[:-1]


def _view_up_matrix(view, up):
    """Matrix that ``pf.Orientations.from_view_up`` hands to scipy's ``Rotation.from_matrix``
    (pyfar, unpinned in the reference's pyproject.toml:32; classes/orientations.py): rows
    are view, up and right = view x up."""
    view, up = np.asarray(view, float), np.asarray(up, float)
    if not np.linalg.norm(view) or not np.linalg.norm(up):
        raise ValueError("View and Up Vectors must have a length.")
    if not np.isclose(0.0, float(np.dot(view, up))):
        raise ValueError("View and Up vectors must be perpendicular.")
    return np.stack([view, up, np.cross(view, up)])


def wall_rotation_euler(wall_normal, wall_up):
    """The 'xyz' Euler angles (degrees) that ``_rotate_coords_to_normal`` (reference
    RadiosityFast.py:971-986) derives for a wall:
    ``(from_view_up(normal, up).inv() * from_view_up([0,0,1], [1,0,0])).as_euler('xyz', True)``
    -- the same scipy calls pyfar's Orientations (a scipy Rotation subclass) makes."""
    import warnings
    from scipy.spatial.transform import Rotation
    o1 = Rotation.from_matrix(_view_up_matrix(wall_normal, wall_up))
    o2 = Rotation.from_matrix(_view_up_matrix([0.0, 0.0, 1.0], [1.0, 0.0, 0.0]))
    with warnings.catch_warnings():             # gimbal lock: scipy picks the third angle = 0
        warnings.simplefilter("ignore", UserWarning)
        return (o1.inv() * o2).as_euler('xyz', True).flatten()


def rotate_to_wall(coords, wall_normal, wall_up):
    """``coords.copy(); .rotate('xyz', euler); .radius = 1`` of the reference
    (RadiosityFast.py:979-985) without pyfar: scipy applies the rotation, the radius is set
    through pyfar's spherical round trip (``cart2sph`` / ``sph2cart`` of
    classes/coordinates.py, which also flushes components below machine epsilon to zero)."""
    from scipy.spatial.transform import Rotation
    rot = Rotation.from_euler('xyz', wall_rotation_euler(wall_normal, wall_up), degrees=True)
    shape = coords.cshape
    pts = rot.apply(np.asarray(coords.cartesian, float).reshape(-1, 3))
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    radius = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    z_div_r = np.divide(z, radius, out=np.zeros_like(radius), where=radius != 0)
    colatitude = np.arccos(z_div_r)
    azimuth = np.mod(np.arctan2(y, x), 2 * np.pi)
    r_sin_cola = 1.0 * np.sin(colatitude)
    out = np.stack([r_sin_cola * np.cos(azimuth), r_sin_cola * np.sin(azimuth),
                    1.0 * np.cos(colatitude)], axis=-1)
    out[np.abs(out) < np.finfo(float).eps] = 0
    res = Coordinates.from_cartesian(out.reshape(shape + (3,)),
                                     weights=getattr(coords, "weights", None))
    return res


def rotation_to_wall_frame(wall_normal, wall_up):
    """Rotation taking the BRDF frame (normal +z, up +x) to a wall's frame.

    The closed form of what :func:`rotate_to_wall` applies (columns = images of the
    BRDF-frame axes x (up), y, z (normal)); kept for tests and for building scenes.
    """
    n = np.asarray(wall_normal, float)
    n = n / np.linalg.norm(n)
    u = np.asarray(wall_up, float)
    u = u - np.dot(u, n) * n
    u = u / np.linalg.norm(u)
    r = np.cross(n, u)
    # columns are the images of the BRDF-frame axes x (up), y, z (normal)
    return np.stack([u, r, n], axis=1)
