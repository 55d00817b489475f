"""sparrowpy_b200 -- B200-native DirectionalRadiosityFast hot path.

Drop-in surface of the reference package for this path::

    import sparrowpy_b200 as sp
    rad = sp.DirectionalRadiosityFast.from_polygon(walls, patch_size)
    rad.bake_geometry(); rad.init_source_energy(src)
    rad.calculate_energy_exchange(343.2, 1e-3, 1.0, max_reflection_order=20)
    etc = rad.collect_energy_receiver_mono(receivers)
"""
__version__ = "0.1.0"

from . import brdf, geometry, pyfar_shim, scenes, sound_object  # noqa: F401
from .geometry import Polygon  # noqa: F401
from .radiosity import DirectionalRadiosityFast  # noqa: F401
from . import testing  # noqa: F401

__all__ = ["DirectionalRadiosityFast", "Polygon", "geometry", "sound_object", "brdf",
           "scenes", "pyfar_shim", "testing"]
