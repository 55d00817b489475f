"""Visibility kernels vs the CPU oracle on scenes the golden fixtures do not cover:
tilted (non-axis-aligned) blocking rectangles, lattice-degenerate grids with awkward
patch sizes, blockers != patches, batched evaluation points.  Bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def random_rect(rng, size=(0.4, 1.6), spread=2.5):
    a = rng.normal(size=3)
    a /= np.linalg.norm(a)
    b = rng.normal(size=3)
    b -= np.dot(a, b) * a
    b /= np.linalg.norm(b)
    o = rng.uniform(-spread, spread, size=3)
    la, lb = rng.uniform(*size, size=2)
    pts = np.array([o, o + la * a, o + la * a + lb * b, o + lb * b])
    return pts, np.cross(a, b)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_tilted_rectangles_p2p(oracle, seed):
    """General orientation: exercises the x87 rotation path, the non-axis-aligned branch
    of the clearance rule and real occlusion between random rectangles."""
    from sparrowpy_b200 import bake
    rng = np.random.default_rng(seed)
    n = 90
    rects = [random_rect(rng) for _ in range(n)]
    pts = np.array([r[0] for r in rects])
    nrm = np.array([r[1] for r in rects])
    cen = pts.mean(axis=1)
    ref = oracle.visibility_p2p(cen, nrm, pts)
    vis = bake.visibility_p2p(T(cen), T(nrm), T(pts)).cpu().numpy()
    assert np.array_equal(vis, ref)
    assert 0 < ref.sum() < n * (n - 1) // 2          # some blocked, some visible


def test_points_against_tilted_walls(oracle):
    from sparrowpy_b200 import bake
    rng = np.random.default_rng(7)
    walls = [random_rect(rng, size=(1.0, 3.0)) for _ in range(25)]
    wp = np.array([w[0] for w in walls])
    wn = np.array([w[1] for w in walls])
    targets = rng.uniform(-3, 3, size=(300, 3))
    points = rng.uniform(-3, 3, size=(7, 3))
    # put some evaluation points and targets exactly on walls
    points[0] = wp[3].mean(axis=0)
    targets[:25] = wp.mean(axis=1)
    vis = bake.visibility_pt2p(T(points), T(targets), T(wn), T(wp)).cpu().numpy()
    for r, p in enumerate(points):
        ref = oracle.visibility_pt2p(p, targets, wn, wp)
        assert np.array_equal(vis[r], ref), r


@pytest.mark.parametrize("patch", [1.0 / 3.0, 0.3, 0.25])
def test_lattice_degenerate_boxes(oracle, patch):
    """Segments through shared edges / vertices of a lattice (the common degenerate
    case of synthetic scenes) with patch sizes that are not exact in binary."""
    from sparrowpy_b200 import bake, geometry, scenes
    walls = scenes.occluder_scene(3, 1, 1) + scenes.building(0, 0, 1, 1, 2)
    wp = np.array([w[0] for w in walls])
    wn = np.array([w[2] for w in walls])
    pts, ids = geometry.process_patches(wp, patch)
    cen = geometry.calculate_center(pts)
    nrm = wn[ids]
    ref = oracle.visibility_p2p(cen, nrm, pts)
    vis = bake.visibility_p2p(T(cen), T(nrm), T(pts)).cpu().numpy()
    assert np.array_equal(vis, ref)
    # blockers = walls (the source/receiver case), evaluation points on the lattice
    grid = np.array([[x, y, z] for x in (0.0, 1.0, 1.5) for y in (0.5, 1.0, 2.0)
                     for z in (0.0, 0.5, 1.0)])
    vis_pt = bake.visibility_pt2p(T(grid), T(cen), T(wn), T(wp)).cpu().numpy()
    for r, p in enumerate(grid):
        assert np.array_equal(vis_pt[r], oracle.visibility_pt2p(p, cen, wn, wp)), r
