"""BRDF builders from scattering coefficients (reference sparrowpy/brdf.py:8-224).

Host-side numpy producing ``FrequencyData`` of shape (n_sources, n_receivers, n_bins)
for ``DirectionalRadiosityFast.set_wall_brdf``.  Writing SOFA files (``file_path``)
needs the sofar package and is not part of the B200 path.
"""
import numpy as np

from . import pyfar_shim

try:
    import pyfar as _pf
    _COORD_TYPES = (pyfar_shim.Coordinates, _pf.Coordinates)
    _FREQ_TYPES = (pyfar_shim.FrequencyData, _pf.FrequencyData)
    _FrequencyData = _pf.FrequencyData
except Exception:  # noqa: BLE001
    _COORD_TYPES = (pyfar_shim.Coordinates,)
    _FREQ_TYPES = (pyfar_shim.FrequencyData,)
    _FrequencyData = pyfar_shim.FrequencyData


def _no_sofa(file_path):
    if file_path is not None:
        raise NotImplementedError(
            "writing SOFA files needs the sofar package and is outside the B200 hot path")


def create_from_scattering(source_directions, receiver_directions,
                           scattering_coefficient, absorption_coefficient=None,
                           file_path=None):
    r"""BRDF of a surface with random-incidence scattering coefficient ``s`` and
    absorption ``alpha`` (reference brdf.py:8-130):

    .. math::
        \rho(\Omega_i, \Omega_o) = \frac{(1-s)(1-\alpha)}{\Omega_i \cdot n}
        \frac{1}{w_o} \delta(\Omega_i - M(\Omega_o)) + \frac{s(1-\alpha)}{\pi}
    """
    if (not isinstance(scattering_coefficient, _FREQ_TYPES)
            or not scattering_coefficient.cshape == (1,)):
        raise TypeError(
            'scattering_coefficient must be a pf.FrequencyData object'
            'with shape (1,)')
    if not isinstance(source_directions, _COORD_TYPES):
        raise TypeError('source_directions must be a pf.Coordinates object')
    if not isinstance(receiver_directions, _COORD_TYPES):
        raise TypeError('receiver_directions must be a pf.Coordinates object')
    _no_sofa(file_path)
    freqs = scattering_coefficient.frequencies
    if absorption_coefficient is None:
        absorption = np.zeros(len(freqs))
    else:
        absorption = np.real(np.asarray(absorption_coefficient.freq)).flatten()
    n_bins = len(freqs)
    brdf = np.zeros((source_directions.csize, receiver_directions.csize, n_bins))
    # in place, as in the reference (brdf.py:103-104): the caller's weights come back
    # normalised to the hemisphere (sum = 2 pi)
    receiver_weights = receiver_directions.weights
    receiver_weights *= 2 * np.pi / np.sum(receiver_weights)
    scattering = np.real(np.asarray(scattering_coefficient.freq)).flatten()
    image_source = source_directions.copy()
    image_source.azimuth = image_source.azimuth + np.pi
    i_receiver = np.asarray(receiver_directions.find_nearest(image_source)[0][0])
    cos_factor = (np.cos(np.asarray(source_directions.colatitude)[np.newaxis, ...])
                  * receiver_weights[..., np.newaxis])
    brdf[:, :, :] += scattering / np.pi
    i_sources = np.arange(source_directions.csize)
    brdf[i_sources, i_receiver, :] += (1 - scattering[np.newaxis, ...]) / cos_factor[
        i_sources, i_receiver, np.newaxis]
    brdf *= (1 - absorption)
    return _FrequencyData(brdf, freqs)


def create_from_directional_scattering(source_directions, receiver_directions,
                                       directional_scattering,
                                       absorption_coefficient=None, file_path=None):
    r"""BRDF from directional scattering coefficients (reference brdf.py:133-224):

    .. math::
        \rho(\Omega_i, \Omega_o) = \frac{(1-\alpha)}{(\Omega_o \cdot n) \, w_o}
        s_d(\Omega_i, \Omega_o)
    """
    if not isinstance(source_directions, _COORD_TYPES):
        raise TypeError('source_directions must be a pf.Coordinates object')
    if not isinstance(receiver_directions, _COORD_TYPES):
        raise TypeError('receiver_directions must be a pf.Coordinates object')
    if (not isinstance(directional_scattering, _FREQ_TYPES)
            or not directional_scattering.cshape == (
                source_directions.csize, receiver_directions.csize)):
        raise TypeError(
            'directional_scattering must be a pf.FrequencyData object with'
            f' cshape ({source_directions.csize, receiver_directions.csize})')
    _no_sofa(file_path)
    freqs = directional_scattering.frequencies
    if absorption_coefficient is None:
        absorption = np.zeros(len(freqs))
    else:
        absorption = np.real(np.asarray(absorption_coefficient.freq)).flatten()
    cos_receiver = np.cos(np.asarray(receiver_directions.colatitude))[
        np.newaxis, :, np.newaxis]
    receiver_weights = receiver_directions.weights        # in place (brdf.py:205-206)
    receiver_weights *= 2 * np.pi / np.sum(receiver_weights)
    brdf = np.real(np.asarray(directional_scattering.freq)) / receiver_weights[
        ..., np.newaxis] / cos_receiver
    brdf = brdf * (1 - absorption)
    return _FrequencyData(brdf, freqs)
