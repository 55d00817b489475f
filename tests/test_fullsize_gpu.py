"""Parity at BASELINE.json's full config-2 size (shoebox 5x6x4 m at 0.2 m, N = 3700,
P = 5.67 M pairs, D = 16, B = 6): the CPU oracle cannot bake this scene in reasonable
time, so the bake is checked on sampled rows / pairs (bit-exact booleans and integers)
and the exchange through size-independent properties."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import bench
    rad = bench.build_scene(bench.CONFIGS["c2"], "f64")
    torch.cuda.synchronize()
    return rad


def test_c2_sizes(c2):
    assert c2.n_patches == 3700
    pairs = c2._baked["pairs"]
    # convex room: every pair of patches on different walls is mutually visible
    wall = c2._patch_to_wall_ids
    counts = np.bincount(wall)
    expected = (3700 ** 2 - int((counts ** 2).sum())) // 2
    assert pairs.shape[0] == expected == 5672500


def test_c2_visibility_rows_bit_exact(c2, oracle):
    rows = [0, 1, 611, 1999, 2500, 3650]
    cen, nrm, pts = c2.patches_center, c2.patches_normal, c2.patches_points
    vis = c2._baked["vis"]
    for r in rows:
        ref = oracle.visibility_p2p(cen, nrm, pts, row_lo=r, row_hi=r + 1)[0]
        assert np.array_equal(vis[r].cpu().numpy(), ref), r


def test_c2_pair_tables_sampled(c2, oracle):
    import bench
    b = c2._baked
    rng = np.random.default_rng(0)
    pairs = b["pairs"].cpu().numpy()
    sel = np.sort(rng.choice(len(pairs), 4000, replace=False))
    # include vertex-sharing (Nusselt) pairs: patches at wall junctions
    ff_ref = oracle.ff_pairs(c2.patches_points, c2.patches_normal, c2.patches_area, pairs[sel])
    ff = b["ff"].cpu().numpy()[sel]
    assert np.max(np.abs(ff - ff_ref) / ff_ref) < 1e-6
    vi = np.array([s.cartesian for s in c2._brdf_incoming_directions])
    vo = np.array([s.cartesian for s in c2._brdf_outgoing_directions])
    brdf = np.array([np.real(x).reshape(vi.shape[1], vo.shape[1], -1) for x in c2._brdf])
    tilde, odir, idir, delay = oracle.pair_tables(
        c2.patches_center, c2.patches_area, c2._patch_to_wall_ids, pairs[sel], ff_ref,
        np.real(c2._air_attenuation), vi, vo, brdf, np.asarray(c2._brdf_index),
        bench.SPEED_OF_SOUND, bench.DT)
    dsel = np.stack([2 * sel, 2 * sel + 1], 1).reshape(-1)
    assert np.array_equal(b["out_dir"].cpu().numpy()[dsel], odir)
    assert np.array_equal(b["in_dir"].cpu().numpy()[dsel], idir)
    from sparrowpy_b200 import bake
    dl = bake.delay_bins(b["dist"], bench.SPEED_OF_SOUND, bench.DT).cpu().numpy()
    assert np.array_equal(np.repeat(dl[sel], 2), delay)
    # dense tilde entries implied by the factored tables
    coef = b["coef"].cpu().numpy()
    mine = b["ff_dir"].cpu().numpy()[dsel, None, None] * coef[b["cls"].cpu().numpy()[dsel]]
    assert np.max(np.abs(mine - tilde)) / np.max(tilde) < 1e-6


def test_c2_form_factor_row_sums(c2):
    """closed room: sum_j F_ij = 1 (reference tests/test_universal_formfactor.py:140-157
    uses 1e-2)."""
    b = c2._baked
    n = c2.n_patches
    sums = torch.zeros(n, dtype=torch.float64, device=b["ff_dir"].device)
    sums.index_add_(0, b["sender"], b["ff_dir"])
    assert float((sums - 1).abs().max()) < 1e-2


def test_c2_exchange_properties(c2):
    import bench
    from sparrowpy_b200 import bake, exchange
    n_samples, c, dt = 1000, bench.SPEED_OF_SOUND, bench.DT
    tables = c2._pair_tables(c, dt, n_samples)
    delay0 = bake.delay_bins(c2._d0_dev, c, dt)
    e0 = c2._e0_dev
    ws = exchange.ExchangeWorkspace(tables, n_samples, e0.device)
    h3 = exchange.energy_exchange(tables, e0, delay0, n_samples, 3, workspace=ws).dense().clone()
    # (1) linearity in the source energy
    h3s = exchange.energy_exchange(tables, 0.25 * e0, delay0, n_samples, 3, workspace=ws).dense()
    assert torch.allclose(h3s, 0.25 * h3, rtol=1e-12, atol=0)
    # (2) superposition of two source energy patterns
    mask = (torch.arange(e0.shape[0], device=e0.device) % 2 == 0)[:, None, None]
    ha = exchange.energy_exchange(tables, e0 * mask, delay0, n_samples, 3, workspace=ws).dense().clone()
    hb = exchange.energy_exchange(tables, e0 * (~mask), delay0, n_samples, 3, workspace=ws).dense()
    assert float((ha + hb - h3).abs().max() / h3.abs().max()) < 1e-12
    # (3) orders add energy monotonically and causally: nothing before the first arrival
    h2 = exchange.energy_exchange(tables, e0, delay0, n_samples, 2, workspace=ws).dense().clone()
    assert float((h3 - h2).min()) >= 0.0
    first = int(delay0.min())
    assert float(h3[..., :first].abs().max()) == 0.0
    # (4) energy of one more order is bounded by the absorbed-and-redistributed previous one
    e2 = float((h2 - exchange.energy_exchange(tables, e0, delay0, n_samples, 1,
                                              workspace=ws).dense()).sum())
    e3 = float((h3 - h2).sum())
    assert 0 < e3 < e2
    # (5) the default (tensor-memory) and the CSR stage 1 agree at full size
    import os
    old = os.environ.get("SPB_GATHER")
    os.environ["SPB_GATHER"] = "csr"
    try:
        hc = exchange.energy_exchange(tables, e0, delay0, n_samples, 3, workspace=ws).dense()
        assert float((hc - h3).abs().max() / h3.abs().max()) < 1e-12
    finally:
        if old is None:
            os.environ.pop("SPB_GATHER", None)
        else:
            os.environ["SPB_GATHER"] = old
