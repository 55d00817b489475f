"""Hierarchical (per-wall) visibility, checked on the CPU: the device predicates of
csrc/vis_group.cuh compiled for the host + the Python group/bin builder reproduce the
oracle's visibility matrix bit for bit (golden scenes, lattice-degenerate scenes,
tilted rectangles, partially invalid groupings)."""
import numpy as np
import pytest

from conftest import load_golden


def grouped(oracle, cen, nrm, pts, ids):
    """Grouped visibility with the result-neutral hints (memoised own-polygon test, own walls
    first) -- checked here to BE result-neutral -- and the oracle's matrix."""
    from sparrowpy_b200 import bake
    vis = bake.visibility_p2p_grouped_host(cen, nrm, pts, ids)
    plain = bake.visibility_p2p_grouped_host(cen, nrm, pts, ids, hints=False)
    assert np.array_equal(vis, plain)
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
    groups, members, bin_ptr, bin_items, strips, group_of = bake.build_groups(blk, ids)
    assert len(groups) == 6 and sorted(members.tolist()) == list(range(len(ids)))
    gi = groups.view(np.int32).reshape(6, -1)
    I = bake._GRP_I
    # every blocker is listed in at least one y-bin and one cell of its own group
    for g in range(6):
        own = set(members[gi[g, I["m0"]]:gi[g, I["m1"]]].tolist())
        b0, nb = gi[g, I["bin_ptr0"]], gi[g, I["n_bins"]]
        assert set(bin_items[bin_ptr[b0]:bin_ptr[b0 + nb]].tolist()) == own
        assert len(own) >= bake._MIN_CELL_MEMBERS and gi[g, I["n_bx"]] > 0
        c0, nc = gi[g, I["cell_ptr0"]], gi[g, I["n_bx"]] * nb
        assert set(bin_items[bin_ptr[c0]:bin_ptr[c0 + nc]].tolist()) == own
        # a cell lists a handful of members, a bin a whole band of the wall
        assert np.diff(bin_ptr[c0:c0 + nc + 1]).max() <= 9
        # strips: sorted, disjoint, two per lattice row boundary at most
        st = strips[gi[g, I["strip0"]]:gi[g, I["strip0"]] + gi[g, I["n_strips"]]]
        assert (st[:, 0] <= st[:, 1]).all() and (st[1:, 0] > st[:-1, 1]).all()
        sf = bin_ptr[gi[g, I["sfirst0"]]:gi[g, I["sfirst0"]] + nb]
        assert (np.diff(sf) >= 0).all() and sf.max() <= len(st)


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


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_tilted_walls_with_cells(oracle, seed):
    """Walls with general normals, 5 x 4 sub-patches each: enough members for the 2-D cells,
    and in the wall's own frame the patches are NOT axis-aligned, so every member is listed in
    all the cells to its left (ray rule only)."""
    from sparrowpy_b200 import bake
    rng = np.random.default_rng(seed)
    pts, nrm, ids = [], [], []
    for w in range(5):
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        b = rng.normal(size=3)
        b -= np.dot(a, b) * a
        b /= np.linalg.norm(b)
        n = np.cross(a, b)
        o = rng.uniform(-1.5, 1.5, size=3)
        la, lb = rng.uniform(0.3, 0.6, size=2)
        for i in range(5):
            for j in range(4):
                p0 = o + i * la * a + j * lb * b
                pts.append([p0, p0 + la * a, p0 + la * a + lb * b, p0 + lb * b])
                nrm.append(n)
                ids.append(w)
    pts, nrm, ids = np.array(pts), np.array(nrm), np.array(ids)
    blk = bake.make_blockers_host(pts, nrm)
    groups = bake.build_groups(blk, ids)[0]
    gi = groups.view(np.int32).reshape(len(groups), -1)
    assert len(groups) == 5 and (gi[:, bake._GRP_I["n_bx"]] > 0).all()
    cen = pts.mean(axis=1)
    # patch centres plus points hovering around the walls: many plane hits inside the lattices
    extra = rng.uniform(-2.5, 2.5, size=(60, 3))
    allc = np.concatenate([cen, extra])
    vis = bake.visibility_p2p_grouped_host(allc, nrm, pts, ids)
    ref = oracle.visibility_p2p(allc, nrm, pts)
    assert np.array_equal(vis, ref)
    assert 0 < ref.sum() < len(allc) * (len(allc) - 1) // 2


def test_cells_equal_bins_on_the_street_canyon():
    """Bench scene C4 (19 200 blockers in 21 walls), 700 sampled patch centres: the cell
    lists give the same matrix as scanning the whole y-bin (cells switched off), which the GPU
    suite holds against the brute-force kernel and the oracle at full size."""
    import ctypes
    from sparrowpy_b200 import _lib, bake, geometry, scenes
    walls = scenes.street_canyon(0, 1.0)
    wp = np.array([w[0] for w in walls])
    wn = np.array([w[2] for w in walls])
    pts, ids = geometry.process_patches(wp, 1.0)
    cen = geometry.calculate_center(pts)
    blk = bake.make_blockers_host(pts, wn[ids])
    groups, members, bin_ptr, bin_items, strips, group_of = bake.build_groups(blk, ids)
    gi = groups.view(np.int32).reshape(len(groups), -1)
    assert (gi[:, bake._GRP_I["n_bx"]] > 0).all()
    sel = np.sort(np.random.default_rng(3).choice(len(cen), 700, replace=False))
    c = np.ascontiguousarray(cen[sel])
    lib = _lib.load()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731

    def run(table):
        vis = np.zeros((len(c), len(c)), np.uint8)
        assert lib.spb_visibility_p2p_grouped_host(
            p(c), ctypes.c_int64(len(c)), p(blk), p(table), ctypes.c_int64(len(table)),
            p(members), p(bin_ptr), p(bin_items), p(strips), ctypes.c_int64(0), None,
            p(vis)) == 0
        return vis

    no_cells = groups.copy()
    no_cells.view(np.int32).reshape(len(groups), -1)[:, bake._GRP_I["n_bx"]] = 0
    new, old = run(groups), run(no_cells)
    assert np.array_equal(new, old)
    assert 0.05 < new.sum() / (len(c) * (len(c) - 1) / 2) < 0.5
