"""Source / receiver value objects (reference sound_object.py:97-252).

``DirectivityMS`` (SOFA-file directivities, reference sound_object.py:8-95) is out of
scope: it needs the sofar package and SOFA fixtures that are not part of this path.
"""
import numpy as np


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
        self.directivity = directivity

    def get_directivity(self, target_position, frequency):
        if self.directivity is None:
            raise ValueError("source has no directivity")
        i_freq = np.argmin(np.abs(self.directivity.data.frequencies - frequency))
        target_position = np.asarray(target_position, float)
        if target_position.size == 3:
            return self.directivity.get_directivity(
                self.position, self.view, self.up, target_position, i_freq)
        return np.array([
            self.directivity.get_directivity(
                self.position, self.view, self.up, pos, i_freq)
            for pos in target_position])[:, 0]


class Receiver(SoundObject):
    """Receiver object."""
