import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu)")


# golden scenes the GPU parity tests run on (all seven fixtures of tests/golden)
GPU_SCENES = ["scene_cube05", "scene_c1", "scene_occluder", "scene_directional",
              "scene_canyon01", "scene_uneven", "scene_canyon015_dir"]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


def rel_err(new, ref):
    """max|new-ref| / max|ref| -- the tolerance metric of BASELINE.md section 3."""
    new, ref = np.asarray(new, float), np.asarray(ref, float)
    scale = np.max(np.abs(ref))
    if scale == 0:
        return float(np.max(np.abs(new)))
    return float(np.max(np.abs(new - ref)) / scale)
