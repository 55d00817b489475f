"""Import the live reference (sparrowpy v1.0.1 at /root/reference) in this container.

Only used by ``make_golden.py`` (fixture generation) -- never at test/bench time:
/root/reference does not exist on the GPU box.  pyfar/sofar/deepdiff/matplotlib are
absent here, so empty module stubs are injected; every numba kernel on the
DirectionalRadiosityFast path runs unchanged on raw ndarrays.
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


def import_reference():
    for name in ["pyfar", "deepdiff", "matplotlib", "matplotlib.axes",
                 "matplotlib.pyplot", "sofar"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)

    class _Dummy:
        pass

    pf = sys.modules["pyfar"]
    for attr in ("Coordinates", "FrequencyData", "TimeData", "Orientations"):
        if not hasattr(pf, attr):
            setattr(pf, attr, _Dummy)
    mpl = sys.modules["matplotlib"]
    mpl.axes = sys.modules["matplotlib.axes"]
    mpl.axes.Axes = _Dummy
    mpl.pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import sparrowpy as sp  # noqa: E402
    from sparrowpy.classes import RadiosityFast as RF
    from sparrowpy import geometry as geo
    from sparrowpy.form_factor import universal as ffu
    from sparrowpy.form_factor import integration as integ
    return sp, RF, geo, ffu, integ
