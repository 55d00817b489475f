"""Scene stubs for tests (reference sparrowpy/testing/stub_utils.py)."""
from . import scenes
from .geometry import Polygon


def shoebox_room_stub(length_x, length_y, length_z):
    """Shoebox room as a list of six ``Polygon`` walls (stub_utils.py:5-48)."""
    return [Polygon(*w) for w in scenes.shoebox(length_x, length_y, length_z)]
