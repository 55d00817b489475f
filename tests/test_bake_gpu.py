"""CUDA bake kernels vs the golden vectors of the live reference and the oracle:
bit-exact booleans / integers, <=1e-6 relative floating point (FP64 path)."""
import numpy as np
import pytest
import torch

from conftest import GPU_SCENES, load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SCENES = GPU_SCENES


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def test_device_x87_norm_is_bit_exact():
    from sparrowpy_b200 import bake
    g = load_golden("rounding_probes")
    for v, ref in ((g["v3"], g["n3"]), (g["v2"], g["n2"]), (g["l3"], g["nl3"])):
        out = bake.probe_norms(T(v)).cpu().numpy()
        assert np.array_equal(out, ref)


def test_basic_visibility_predicates_bit_exact():
    from sparrowpy_b200 import bake
    g = load_golden("predicates")
    vis, in_a, in_b = bake.probe_basic_visibility(T(g["A"]), T(g["B"]), T(g["S"]), T(g["N"]))
    assert np.array_equal(in_a.cpu().numpy(), g["inA"])
    assert np.array_equal(in_b.cpu().numpy(), g["inB"])
    assert np.array_equal(vis.cpu().numpy(), g["visible"])


def test_form_factor_pairs():
    from sparrowpy_b200 import bake
    g = load_golden("form_factor_pairs")
    n = len(g["ff"])
    pts = np.concatenate([g["pts_i"], g["pts_j"]])
    nrm = np.concatenate([g["normal_i"], g["normal_j"]])
    areas = np.concatenate([g["area_i"], np.ones(n)])
    pairs = np.stack([np.arange(n), np.arange(n) + n], 1).astype(np.int32)
    ff, flag = bake.form_factors(T(pts), T(nrm), T(areas), T(pairs))
    assert np.array_equal(flag.cpu().numpy(), g["nusselt"])
    err = np.abs(ff.cpu().numpy() - g["ff"]) / np.abs(g["ff"])
    assert err.max() < 1e-6, err.max()


def test_point_patch_factors(oracle):
    from sparrowpy_b200 import bake
    g = load_golden("point_patch")
    n = len(g["points"])
    # one receiver per call row: use the batched receiver kernel with N = 1 patch each
    for k in range(0, n, 37):
        patch = g["patches"][k:k + 1]
        cen = patch.mean(axis=1)
        out = bake.receiver_factors(
            T(g["points"][k:k + 1]), T(cen), T(patch), T(np.ones((1, 1), np.uint8)),
            T(np.zeros(1)), T(np.zeros(1, np.int64)), T(np.array([[[0.0, 0, 1]]])),
            343.2, 1e-3, 100)
        ref = g["receiver"][k]
        if np.isfinite(ref):
            assert abs(out["factor"].item() - ref) <= 1e-9 * max(1.0, abs(ref))


@pytest.mark.parametrize("name", SCENES)
def test_scene_bake_matches_reference(oracle, name):
    from sparrowpy_b200 import bake, geometry
    g = load_golden(name)
    pts_h, ids_h = geometry.process_patches(g["walls_points"], float(g["patch_size"]))
    assert np.array_equal(pts_h, g["patches_points"])
    cen_h, area_h = geometry.calculate_center(pts_h), geometry.calculate_area(pts_h)
    assert np.array_equal(cen_h, g["patches_center"])
    assert np.array_equal(area_h, g["patches_area"])
    n = len(ids_h)
    nrm_h = g["walls_normal"][ids_h]
    pts, cen, nrm, area = T(pts_h), T(cen_h), T(nrm_h), T(area_h)

    # visibility: bit-exact
    vis = bake.visibility_p2p(cen, nrm, pts)
    vis_ref = np.unpackbits(g["visibility"])[:n * n].reshape(n, n).astype(bool)
    assert np.array_equal(vis.cpu().numpy(), vis_ref)
    pairs = bake.visible_pairs(vis)
    assert pairs.dtype == torch.int32
    assert np.array_equal(pairs.cpu().numpy(), g["visible_patches"])

    # form factors: tolerance
    ff, _ = bake.form_factors(pts, nrm, area, pairs)
    assert rel_err(ff.cpu().numpy(), g["ff_pairs"]) < 1e-6
    assert np.max(np.abs(ff.cpu().numpy() - g["ff_pairs"]) / g["ff_pairs"]) < 1e-6

    # pair tables: bit-exact integers
    c, dt = float(g["speed_of_sound"]), float(g["dt"])
    vi, vo = T(g["vi"]), T(g["vo"])
    baked = "brdf_dirs" in g
    dist, out_dir, in_dir = bake.pair_geometry(cen, T(ids_h), pairs, vi if baked else None,
                                               vo if baked else None)
    delays = bake.delay_bins(dist, c, dt).cpu().numpy()
    assert np.array_equal(delays, g["pair_delays"])
    p_h = g["visible_patches"]
    p2o = g["p2o"].astype(np.int64)
    assert np.array_equal(out_dir.cpu().numpy()[0::2], p2o[p_h[:, 0], p_h[:, 1]])
    assert np.array_equal(out_dir.cpu().numpy()[1::2], p2o[p_h[:, 1], p_h[:, 0]])
    ref = oracle.pair_tables(cen_h, area_h, ids_h, p_h, g["ff_pairs"],
                             g["air_attenuation"], g["vi"], g["vo"],
                             g["brdf"].reshape(g["brdf"].shape[0], g["vi"].shape[1],
                                               g["vo"].shape[1], -1), g["brdf_index"], c, dt)
    if baked:
        assert np.array_equal(in_dir.cpu().numpy(), ref[2])
        assert np.array_equal(out_dir.cpu().numpy(), ref[1])

    # source: visibility + distance bit-exact, energies within tolerance
    wp, wn = T(g["walls_points"]), T(g["walls_normal"])
    svis = bake.visibility_pt2p(T(g["source"]), cen, wn, wp)[0]
    assert np.array_equal(svis.cpu().numpy(), g["source_visibility"])
    brdf = g["brdf"].reshape(g["brdf"].shape[0], g["vi"].shape[1], g["vo"].shape[1], -1)
    d0, e0, en = bake.source_energy(T(g["source"]), cen, pts, svis, T(g["air_attenuation"]),
                                    T(ids_h), vi, T(brdf), T(g["brdf_index"]),
                                    g["vo"].shape[1])
    assert np.array_equal(d0.cpu().numpy(), g["distance_patches_to_source"])
    assert np.array_equal(bake.delay_bins(d0, c, dt).cpu().numpy(), g["source_delays"])
    assert rel_err(en.cpu().numpy(), g["energy_0"]) < 1e-6
    assert rel_err(e0.cpu().numpy(), g["energy_init_source"]) < 1e-6

    # receivers
    n_samples = int(float(g["duration"]) / dt)
    rvis = bake.visibility_pt2p(T(g["receivers"]), cen, wn, wp)
    assert np.array_equal(rvis.cpu().numpy(), g["receiver_visibility"])
    rf = bake.receiver_factors(T(g["receivers"]), cen, pts, rvis, T(g["air_attenuation"]),
                               T(ids_h), vo, c, dt, n_samples)
    assert np.array_equal(rf["rdir"].cpu().numpy(), g["receiver_dir_index"])
    assert np.array_equal(rf["delay"].cpu().numpy(), g["receiver_delays"])
    assert rel_err(rf["factor"].cpu().numpy(), g["receiver_factor"]) < 1e-6


@pytest.mark.parametrize("name", SCENES)
def test_grouped_visibility_equals_bruteforce_and_reference(name):
    """Hierarchical (per-wall) visibility kernel: same matrix as the brute-force kernel
    and as the reference."""
    from sparrowpy_b200 import bake
    g = load_golden(name)
    n = len(g["patches_center"])
    cen, nrm, pts = T(g["patches_center"]), T(g["patches_normal"]), T(g["patches_points"])
    brute = bake.visibility_p2p(cen, nrm, pts)
    grouped = bake.visibility_p2p_grouped(cen, nrm, pts, g["patch_to_wall_ids"])
    gold = np.unpackbits(g["visibility"])[:n * n].reshape(n, n).astype(bool)
    assert torch.equal(brute, grouped)
    assert np.array_equal(grouped.cpu().numpy(), gold)


def test_grouped_visibility_full_size_c2():
    """N = 3700: grouped == brute force on the whole matrix."""
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import bake, geometry, scenes
    walls = scenes.shoebox(5, 6, 4)
    wp = np.array([w[0] for w in walls])
    wn = np.array([w[2] for w in walls])
    pts, ids = geometry.process_patches(wp, 0.2)
    cen = geometry.calculate_center(pts)
    brute = bake.visibility_p2p(T(cen), T(wn[ids]), T(pts))
    grouped = bake.visibility_p2p_grouped(T(cen), T(wn[ids]), T(pts), ids)
    assert torch.equal(brute, grouped)
    assert int(grouped.sum()) == 5672500


@pytest.mark.parametrize("name", ["scene_directional", "scene_canyon01", "scene_uneven"])
@pytest.mark.parametrize("n_parts", [2, 3])
def test_sharded_bake_builds_the_same_tables(name, n_parts):
    """distributed.sharded_bake_tables: the bake split by visibility rows over `n_parts`
    ranks (played by one process; the all-to-all is replaced by an in-memory router) gives
    exactly the tables of the unsharded bake restricted to each receiver shard."""
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import distributed, pyfar_shim as pf
    g = load_golden(name)
    walls = [sp.Polygon(p, u, n) for p, u, n in
             zip(g["walls_points"], g["walls_up"], g["walls_normal"])]
    rad = sp.DirectionalRadiosityFast.from_polygon(walls, float(g["patch_size"]))
    if "brdf_dirs" in g:
        coords = pf.Coordinates.from_cartesian(g["brdf_dirs"], weights=g["brdf_weights"])
        for m in range(g["brdf"].shape[0]):
            rad.set_wall_brdf(np.nonzero(g["brdf_index"] == m)[0],
                              pf.FrequencyData(g["brdf"][m] / np.pi, g["frequencies"]),
                              coords, coords)
        rad.set_air_attenuation(pf.FrequencyData(g["air_attenuation"], g["frequencies"]))
    c, dt = float(g["speed_of_sound"]), float(g["dt"])
    n_samples = int(float(g["duration"]) / dt)
    # pass 1: every part computes its rows and posts its directed pairs
    posted = []

    def collect(dest, ints, ff, stats):
        posted.append((dest.clone(), ints.clone(), ff.clone(), stats.clone()))
        return ints[:0], ff[:0], stats

    for part in range(n_parts):
        distributed.sharded_bake_tables(rad, c, dt, n_samples, part=part, n_parts=n_parts,
                                        route=collect)
    dest = torch.cat([p[0] for p in posted])
    ints = torch.cat([p[1] for p in posted])
    ff = torch.cat([p[2] for p in posted])
    stats = torch.stack([torch.stack([p[3][0] for p in posted]).sum(),
                         torch.stack([p[3][1] for p in posted]).max()])
    # pass 2: every part receives what was addressed to it (in scrambled order)
    gen = torch.Generator(device="cpu").manual_seed(3)
    rad.bake_geometry()
    assert int(stats[0]) == rad._baked["pairs"].shape[0]
    for part in range(n_parts):
        mine = torch.nonzero(dest == part).reshape(-1)
        mine = mine[torch.randperm(mine.numel(), generator=gen).to(mine.device)]

        def deliver(_d, _i, _f, _s, mine=mine):
            return ints[mine], ff[mine], stats

        got, n_pairs = distributed.sharded_bake_tables(rad, c, dt, n_samples, part=part,
                                                       n_parts=n_parts, route=deliver)
        rad._tables = None
        want = rad._pair_tables(c, dt, n_samples, n_shards=n_parts, shard=part)
        assert n_pairs == rad._baked["pairs"].shape[0]
        assert (got.n_patches, got.max_delay, got.n_directed, got.win_w) == \
            (want.n_patches, want.max_delay, want.n_directed, want.win_w)
        for f in ("seg_ptr", "src", "wgt", "dly", "coef", "win_ptr", "win_recs", "rank"):
            assert torch.equal(getattr(got, f), getattr(want, f)), (part, f)
