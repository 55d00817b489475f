"""Source / receiver value objects (reference sound_object.py:97-252) and source
directivities (reference sound_object.py:8-95).

A directivity enters the hot path as one real factor per (patch, band) that scales the
initial energy (RadiosityFast.py:497-517) and one per (receiver, band) for the direct
sound (:648-651); it is evaluated on the host (nearest measured direction) and
multiplied into the device tensors.  ``DirectivityMS(file_path)`` reads SOFA files
through sofar like the reference and therefore needs that package;
``DirectivityMS.from_arrays`` builds the same object from plain arrays.
"""
import numpy as np


def _get_metrics(pos_g, view_g, up_g, target_pos_g):
    """Azimuth and elevation (degrees) of ``target`` seen from a source at ``pos`` that
    looks along ``view`` with ``up`` (reference sound_object.py:67-87): local axes
    x' = view x up, y' = up, z' = -view; azimuth = atan2(-x', -z'), elevation = asin(y'/|w|).
    Accepts one target (3,) or many (n, 3)."""
    pos_g = np.asarray(pos_g, dtype=float)
    view_g = np.asarray(view_g, dtype=float)
    up_g = np.asarray(up_g, dtype=float)
    direction = np.asarray(target_pos_g, dtype=float) - pos_g
    x_dash = np.cross(view_g, up_g)
    w_x = direction @ x_dash
    w_y = direction @ up_g
    w_z = direction @ (-view_g)
    azimuth_deg = np.arctan2(-w_x, -w_z) / np.pi * 180
    elevation_deg = np.arcsin(w_y / np.sqrt(w_x * w_x + w_y * w_y + w_z * w_z)) / np.pi * 180
    return azimuth_deg, elevation_deg


class _Spectrum:
    """Stand-in for the ``pf.FrequencyData`` the reference keeps in ``DirectivityMS.data``
    (only ``.freq`` and ``.frequencies`` are used)."""

    def __init__(self, freq, frequencies):
        self.freq = np.asarray(freq)
        self.frequencies = np.asarray(frequencies, dtype=float)


class DirectivityMS:
    """Directivity in the FreeFieldDirectivityTF convention: complex transfer factors
    ``data.freq (n_directions, n_frequencies)`` measured at ``receivers`` (unit-sphere
    directions in the source frame)."""

    def __init__(self, file_path, source_index=0):
        try:
            import sofar as sf
        except ImportError as exc:      # pragma: no cover - sofar absent in this image
            raise ImportError(
                "DirectivityMS(file_path) reads SOFA files through the sofar package; "
                "use DirectivityMS.from_arrays(data, frequencies, receiver_positions) "
                "when it is not installed") from exc
        sofa = sf.read_sofa(file_path, verbose=False)          # pragma: no cover
        data = sofa.Data_Real[source_index, :] + 1j * sofa.Data_Imag[source_index, :]
        positions = np.squeeze(np.asarray(sofa.ReceiverPosition, float))
        if positions.ndim != 2:
            raise ValueError(
                'DirectivityMS only supports 1D coordinates, '
                f'got {positions.ndim - 1}D coordinates. Squeezing did not work.')
        if str(sofa.ReceiverPosition_Type).lower().startswith("spherical"):
            az, el, rad = np.deg2rad(positions[:, 0]), np.deg2rad(positions[:, 1]), \
                positions[:, 2]
            positions = np.stack([rad * np.cos(el) * np.cos(az),
                                  rad * np.cos(el) * np.sin(az), rad * np.sin(el)], axis=1)
        self._init(data, sofa.N, positions)

    @classmethod
    def from_arrays(cls, data, frequencies, receiver_positions):
        """``data`` (n_directions, n_frequencies) complex or real, ``frequencies``
        (n_frequencies,), ``receiver_positions`` (n_directions, 3) cartesian."""
        self = cls.__new__(cls)
        self._init(data, frequencies, receiver_positions)
        return self

    def _init(self, data, frequencies, positions):
        positions = np.asarray(positions, dtype=float).reshape(-1, 3)
        data = np.asarray(data)
        if data.shape != (positions.shape[0], np.size(frequencies)):
            raise ValueError("data must have shape (n_directions, n_frequencies)")
        self.data = _Spectrum(data, frequencies)
        self.receivers = positions
        self._unit = positions / np.linalg.norm(positions, axis=1)[:, None]

    def nearest_index(self, source_pos, source_view, source_up, target_position):
        """Index of the measured direction nearest to each target, (n,)."""
        target = np.atleast_2d(np.asarray(target_position, float))
        az, el = _get_metrics(source_pos, source_view, source_up, target)
        az, el = np.deg2rad(az), np.deg2rad(el)
        find = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
        out = np.empty(find.shape[0], dtype=np.int64)
        step = max(1, (1 << 24) // max(1, self._unit.shape[0]))
        for k in range(0, find.shape[0], step):      # first minimum, like find_nearest
            d2 = ((find[k:k + step, None, :] - self._unit[None, :, :]) ** 2).sum(-1)
            out[k:k + step] = np.argmin(d2, axis=1)
        return out

    def get_directivity(self, source_pos, source_view, source_up, target_position, i_freq):
        """Nearest directivity factor for one position, shape (1,) like the reference
        (sound_object.py:40-65)."""
        idx = self.nearest_index(source_pos, source_view, source_up, target_position)
        return self.data.freq[idx[:1], i_freq]

    def factors(self, source_pos, source_view, source_up, target_positions, i_freq):
        """The same for many positions at once, shape (n,)."""
        idx = self.nearest_index(source_pos, source_view, source_up, target_positions)
        return self.data.freq[idx, i_freq]


class SoundObject:
    """Position, view and up vector of a source or receiver."""

    def __init__(self, position, view, up):
        self.position = np.array(position, dtype=float)
        assert self.position.shape == (3,)
        self.view = np.array(view, dtype=float)
        self.view /= np.sqrt(np.dot(view, view))
        assert self.view.shape == (3,)
        self.up = np.array(up, dtype=float)
        self.up /= np.sqrt(np.dot(up, up))
        assert self.up.shape == (3,)


class SoundSource(SoundObject):
    """Acoustic sound source (omnidirectional unless a directivity is given)."""

    def __init__(self, position, view, up, directivity=None, sound_power=1):
        super().__init__(position, view, up)
        self.sound_power = float(sound_power)
        if directivity is not None:
            assert isinstance(directivity, DirectivityMS)
        self.directivity = directivity

    def get_directivity(self, target_position, frequency):
        """Nearest directivity factor for position(s) (3,) or (n, 3) and a frequency in
        Hz (reference sound_object.py:193-218)."""
        if self.directivity is None:
            raise ValueError("source has no directivity")
        i_freq = np.argmin(np.abs(self.directivity.data.frequencies - frequency))
        target_position = np.asarray(target_position, float)
        if target_position.size == 3:
            return self.directivity.get_directivity(
                self.position, self.view, self.up, target_position, i_freq)
        return self.directivity.factors(self.position, self.view, self.up,
                                        target_position.reshape(-1, 3), i_freq)


class Receiver(SoundObject):
    """Receiver object."""
