"""Parity at the size of the headline bench workload, BASELINE.json config 4 (street canyon,
N = 19 200 patches, P = 22.3 M visible pairs, occluding buildings, diffuse, T = 2000).  The
CPU oracle cannot bake this scene in reasonable time (O(N^3) visibility), so the bake is
checked on sampled rows / pairs (bit-exact booleans and integers, form factors <= 1e-6) and
the energy exchange against the oracle's `_energy_exchange` on the FULL pair list for two
reflection orders (bounded CPU time), plus the kernels against each other."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c4():
    import bench
    rad = bench.build_scene(bench.CONFIGS["c4"], "f64")
    torch.cuda.synchronize()
    return rad


def test_c4_sizes(c4):
    assert c4.n_patches == 19200
    pairs = c4._baked["pairs"]
    assert pairs.shape[0] == 22291920
    vis = c4._baked["vis"]
    # upper triangle only (geometry.py:771-779), and the buildings do occlude: far fewer
    # pairs than in a convex scene with the same walls
    assert int(torch.tril(vis).sum()) == 0
    wall = c4._patch_to_wall_ids
    counts = np.bincount(wall)
    convex = (19200 ** 2 - int((counts ** 2).sum())) // 2
    assert pairs.shape[0] < 0.2 * convex


def test_c4_visibility_rows_bit_exact(c4, oracle):
    """Rows of the visibility matrix against the oracle's literal O(N) blocker loop
    (geometry.py:750-797): ground patches next to and far from buildings, facade and roof
    patches (the occluded cases), first and late rows."""
    wall = c4._patch_to_wall_ids
    rng = np.random.default_rng(4)
    rows = [0, 1, 7199, int(np.nonzero(wall == 1)[0][0]), int(np.nonzero(wall == 3)[0][5]),
            int(np.nonzero(wall == wall.max())[0][0])]
    rows += [int(r) for r in rng.choice(c4.n_patches - 2000, 4, replace=False)]
    cen, nrm, pts = c4.patches_center, c4.patches_normal, c4.patches_points
    vis = c4._baked["vis"]
    n_true = 0
    for r in sorted(set(rows)):
        ref = oracle.visibility_p2p(cen, nrm, pts, row_lo=r, row_hi=r + 1)[0]
        got = vis[r].cpu().numpy().astype(bool)
        assert np.array_equal(got, ref), r
        n_true += int(ref.sum())
    assert n_true > 0


def test_c4_pair_tables_sampled(c4, oracle):
    import bench
    from sparrowpy_b200 import bake
    b = c4._baked
    rng = np.random.default_rng(1)
    pairs = b["pairs"].cpu().numpy()
    sel = np.sort(rng.choice(len(pairs), 6000, replace=False))
    ff_ref = oracle.ff_pairs(c4.patches_points, c4.patches_normal, c4.patches_area, pairs[sel])
    ff = b["ff"].cpu().numpy()[sel]
    assert np.max(np.abs(ff - ff_ref) / ff_ref) < 1e-6
    vi = np.array([s.cartesian for s in c4._brdf_incoming_directions])
    vo = np.array([s.cartesian for s in c4._brdf_outgoing_directions])
    brdf = np.array([np.real(x).reshape(vi.shape[1], vo.shape[1], -1) for x in c4._brdf])
    tilde, odir, idir, delay = oracle.pair_tables(
        c4.patches_center, c4.patches_area, c4._patch_to_wall_ids, pairs[sel], ff_ref,
        np.real(c4._air_attenuation), vi, vo, brdf, np.asarray(c4._brdf_index),
        bench.SPEED_OF_SOUND, bench.DT)
    dsel = np.stack([2 * sel, 2 * sel + 1], 1).reshape(-1)
    assert np.array_equal(b["out_dir"].cpu().numpy()[dsel], odir)
    assert np.array_equal(b["in_dir"].cpu().numpy()[dsel], idir)
    dl = bake.delay_bins(b["dist"], bench.SPEED_OF_SOUND, bench.DT).cpu().numpy()
    assert np.array_equal(np.repeat(dl[sel], 2), delay)          # delay bins bit-exact
    coef = b["coef"].cpu().numpy()
    mine = b["ff_dir"].cpu().numpy()[dsel, None, None] * coef[b["cls"].cpu().numpy()[dsel]]
    assert np.max(np.abs(mine - tilde)) / np.max(tilde) < 1e-6


def test_c4_source_visibility_and_energy(c4, oracle):
    """Source -> patch visibility (walls as blockers, geometry.py:799-839) bit-exact and the
    initial energy / distances of every patch against the oracle."""
    import bench
    src = np.array(bench.CONFIGS["c4"]["source"], float)
    ref = oracle.visibility_pt2p(src, c4.patches_center, c4.walls_normal, c4.walls_points)
    assert np.array_equal(c4._source_visibility.astype(bool), ref)
    assert 0 < ref.sum() < c4.n_patches                          # some patches are shadowed
    energy, dist = oracle.source_energy(src, c4.patches_center, c4.patches_points, ref,
                                        np.real(c4._air_attenuation))
    assert np.array_equal(c4._distance_patches_to_source, dist)   # feeds the delay bins
    e0 = c4._energy_init_source                                    # (N, 1, 1), diffuse BRDF x pi
    vi = np.array([s.cartesian for s in c4._brdf_incoming_directions])
    vo = np.array([s.cartesian for s in c4._brdf_outgoing_directions])
    brdf = np.array([np.real(x).reshape(vi.shape[1], vo.shape[1], -1) for x in c4._brdf])
    want = oracle.add_directional(energy, src, c4.patches_center, c4._patch_to_wall_ids, vi,
                                  vo, brdf, np.asarray(c4._brdf_index))
    assert np.max(np.abs(e0 - want)) / np.max(want) < 1e-6
    assert np.array_equal(e0 == 0, want == 0)


def test_c4_exchange_two_orders_vs_oracle(c4, oracle):
    """`_energy_exchange` (RadiosityFast.py:1073-1145) on the full 22.3 M pair list, orders
    0..2, against the oracle on all host threads: ETC <= 1e-6 (f64 tolerance of north_star)
    with an identical zero pattern (= every delay bin bit-exact)."""
    import os
    import bench
    from sparrowpy_b200 import bake, exchange
    b = c4._baked
    n_samples, c, dt = 2000, bench.SPEED_OF_SOUND, bench.DT
    tables = c4._pair_tables(c, dt, n_samples)
    delay0 = bake.delay_bins(c4._d0_dev, c, dt)
    got = exchange.energy_exchange(tables, c4._e0_dev, delay0, n_samples, 2).dense()
    pairs = b["pairs"].cpu().numpy()
    tilde = (b["ff_dir"][:, None, None] * b["coef"][b["cls"]]).cpu().numpy()
    delay = np.repeat((b["dist"].cpu().numpy() / c / dt).astype(np.int64), 2)
    ref = oracle.energy_exchange(
        c4._energy_init_source, c4._distance_patches_to_source, pairs, tilde,
        b["out_dir"].cpu().numpy().astype(np.int64), delay, n_samples, c, dt, 2,
        n_threads=max(1, len(os.sched_getaffinity(0))))
    ref_t = torch.from_numpy(ref).to(got.device)
    err = float((got - ref_t).abs().max() / ref_t.abs().max())
    assert err < 1e-6, err
    assert bool(torch.equal(got == 0, ref_t == 0))
    # sampled rows to a much tighter bound: the kernels sum in the reference's pair order
    rows = torch.tensor([0, 17, 5000, 12345, 19199], device=got.device)
    scale = ref_t[rows].abs().amax()
    assert float((got[rows] - ref_t[rows]).abs().max() / scale) < 1e-12


def test_c4_gather_kernels_agree(c4):
    """Tensor-memory, TMA-tiled and CSR stage 1 give the same histogram at full size."""
    import os
    import bench
    from sparrowpy_b200 import bake, exchange
    n_samples, c, dt = 2000, bench.SPEED_OF_SOUND, bench.DT
    delay0 = bake.delay_bins(c4._d0_dev, c, dt)
    out = {}
    old = os.environ.get("SPB_GATHER")
    try:
        for kind in ("tmem", "tma", "csr"):
            os.environ["SPB_GATHER"] = kind
            c4._tables = None
            tables = c4._pair_tables(c, dt, n_samples)
            out[kind] = exchange.energy_exchange(tables, c4._e0_dev, delay0, n_samples,
                                                 3).dense().clone()
            del tables
    finally:
        c4._tables = None
        if old is None:
            os.environ.pop("SPB_GATHER", None)
        else:
            os.environ["SPB_GATHER"] = old
    scale = out["csr"].abs().max()
    assert float(scale) > 0
    for kind in ("tmem", "tma"):
        assert float((out[kind] - out["csr"]).abs().max() / scale) < 1e-12, kind
    # C4 is diffuse: the default runs the whole order in one kernel; unfused = bit-identical
    os.environ["SPB_FUSED_ORDER"] = "0"
    try:
        tables = c4._pair_tables(c, dt, n_samples)
        unfused = exchange.energy_exchange(tables, c4._e0_dev, delay0, n_samples, 3).dense()
        assert torch.equal(unfused, out["tmem"])
    finally:
        os.environ.pop("SPB_FUSED_ORDER", None)
