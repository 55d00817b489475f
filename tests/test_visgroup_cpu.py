"""Hierarchical (per-wall) visibility, checked on the CPU: the device predicates of
csrc/vis_group.cuh compiled for the host + the Python group/bin builder reproduce the
oracle's visibility matrix bit for bit (golden scenes, lattice-degenerate scenes,
tilted rectangles, partially invalid groupings)."""
import numpy as np
import pytest

from conftest import load_golden


def grouped(oracle, cen, nrm, pts, ids):
    from sparrowpy_b200 import bake
    vis = bake.visibility_p2p_grouped_host(cen, nrm, pts, ids)
    ref = oracle.visibility_p2p(cen, nrm, pts)
    return vis, ref


@pytest.mark.parametrize("name", ["scene_cube05", "scene_c1", "scene_occluder",
                                  "scene_directional", "scene_canyon01", "scene_uneven",
                                  "scene_canyon015_dir"])
def test_golden_scenes(oracle, name):
    g = load_golden(name)
    n = len(g["patches_center"])
    vis, ref = grouped(oracle, g["patches_center"], g["patches_normal"], g["patches_points"],
                       g["patch_to_wall_ids"])
    gold = np.unpackbits(g["visibility"])[:n * n].reshape(n, n).astype(bool)
    assert np.array_equal(ref, gold)
    assert np.array_equal(vis, gold)


@pytest.mark.parametrize("patch", [1.0 / 3.0, 0.3, 0.25, 0.5])
def test_lattice_degenerate(oracle, patch):
    from sparrowpy_b200 import geometry, scenes
    walls = scenes.occluder_scene(3, 1, 1) + scenes.building(0, 0, 1, 1, 2)
    wp = np.array([w[0] for w in walls])
    wn = np.array([w[2] for w in walls])
    pts, ids = geometry.process_patches(wp, patch)
    cen = geometry.calculate_center(pts)
    vis, ref = grouped(oracle, cen, wn[ids], pts, ids)
    assert np.array_equal(vis, ref)
    assert 0 < ref.sum() < ref.size // 2


def test_group_table_layout():
    from sparrowpy_b200 import bake, geometry, scenes
    walls = scenes.shoebox(2, 3, 2)
    wp = np.array([w[0] for w in walls])
    wn = np.array([w[2] for w in walls])
    pts, ids = geometry.process_patches(wp, 0.5)
    blk = bake.make_blockers_host(pts, wn[ids])
    assert blk.shape == (len(ids), 38) and (blk[:, 37] == 1.0).all()      # all axis-aligned
    groups, members, bin_ptr, bin_items = bake.build_groups(blk, ids)
    assert len(groups) == 6 and sorted(members.tolist()) == list(range(len(ids)))
    gi = groups.view(np.int32).reshape(6, -1)
    assert gi[:, 32].sum() == len(bin_ptr) - 1                             # bins of all groups
    # every blocker is listed in at least one bin of its own group
    for g in range(6):
        b0, nb = gi[g, 33], gi[g, 32]
        listed = set(bin_items[bin_ptr[b0]:bin_ptr[b0 + nb]].tolist())
        assert listed == set(members[gi[g, 34]:gi[g, 35]].tolist())


@pytest.mark.parametrize("seed", [0, 1])
def test_tilted_rectangles_and_mixed_groups(oracle, seed):
    """Tilted rectangles tessellated into coplanar sub-patches (general normals: the
    ray-band candidates extend to the left), plus deliberately wrong wall ids that must
    fall back to singleton groups."""
    rng = np.random.default_rng(seed)
    pts, nrm, ids = [], [], []
    for w in range(7):
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        b = rng.normal(size=3)
        b -= np.dot(a, b) * a
        b /= np.linalg.norm(b)
        n = np.cross(a, b)
        o = rng.uniform(-2, 2, size=3)
        la, lb = rng.uniform(0.4, 0.9, size=2)
        for i in range(3):
            for j in range(2):
                p0 = o + i * la * a + j * lb * b
                pts.append([p0, p0 + la * a, p0 + la * a + lb * b, p0 + lb * b])
                nrm.append(n)
                ids.append(w if w != 6 else 5)          # wall 6 is mislabelled as wall 5
    pts, nrm, ids = np.array(pts), np.array(nrm), np.array(ids)
    cen = pts.mean(axis=1)
    vis, ref = grouped(oracle, cen, nrm, pts, ids)
    assert np.array_equal(vis, ref)
    assert 0 < ref.sum() < len(cen) * (len(cen) - 1) // 2
