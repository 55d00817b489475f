"""CUDA energy exchange / collection vs the CPU oracle and the golden fixtures.

The baked inputs (pairs, form factors, direction indices, delays, E0) come from
the oracle here, so the exchange kernels are validated in isolation; the bake
kernels have their own parity tests."""
import numpy as np
import pytest
import torch

from conftest import GPU_SCENES, load_golden, rel_err

pytestmark = pytest.mark.gpu

SCENES = GPU_SCENES
TOL = {"f64": 1e-6, "f32": 1e-4}   # BASELINE.json north_star tolerances


def oracle_run(oracle, g):
    return oracle.pipeline(
        g["walls_points"], g["walls_normal"], float(g["patch_size"]), g["source"],
        g["receivers"], float(g["speed_of_sound"]), float(g["dt"]),
        float(g["duration"]), int(g["max_order"]), g["air_attenuation"], g["vi"],
        g["vo"], g["brdf"].reshape(g["brdf"].shape[0], g["vi"].shape[1],
                                   g["vo"].shape[1], -1), g["brdf_index"],
        brdf_set_before_bake="brdf_dirs" in g)


def device_tables(g, out, dtype, n_samples):
    from sparrowpy_b200 import exchange
    dev = torch.device("cuda:0")
    baked = "brdf_dirs" in g
    n = out["patches_center"].shape[0]
    pairs = torch.from_numpy(out["visible_patches"]).to(dev)
    ff = torch.from_numpy(out["ff_pairs"]).to(dev)
    areas = torch.from_numpy(out["patches_area"]).to(dev)
    sender, receiver, ffd = exchange.directed_pairs(pairs, ff, areas)
    delay = torch.from_numpy(out["pair_delays"]).to(dev)
    out_dir = torch.from_numpy(out["out_dir"]).to(dev)
    if baked:
        s_in = g["vi"].shape[1]
        brdf = g["brdf"].reshape(g["brdf"].shape[0], s_in, g["vo"].shape[1], -1)
        wall = torch.from_numpy(out["patch_to_wall_ids"]).to(dev)
        bidx = torch.from_numpy(g["brdf_index"]).to(dev)
        cls = bidx[wall[sender]] * s_in + torch.from_numpy(out["in_dir"]).to(dev)
        coef = np.exp(-g["air_attenuation"])[None, None, :] * brdf.reshape(
            -1, brdf.shape[2], brdf.shape[3])
    else:
        cls = torch.zeros_like(sender)
        coef = np.ones((1, 1, 1))
    coef = torch.from_numpy(np.ascontiguousarray(coef)).to(dev)
    return exchange.build_pair_tables(sender, receiver, ffd, delay, out_dir, cls, coef,
                                      n, n_samples, dtype)


@pytest.mark.parametrize("gather", ["tmem", "tma", "csr"])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("name", SCENES)
def test_exchange_matches_oracle(oracle, name, dtype, gather, monkeypatch):
    from sparrowpy_b200 import exchange
    monkeypatch.setenv("SPB_GATHER", gather)
    g = load_golden(name)
    out = oracle_run(oracle, g)
    etc_ref = out["etc"]
    n_samples = etc_ref.shape[-1]
    dev = torch.device("cuda:0")
    tables = device_tables(g, out, dtype, n_samples)
    e0 = torch.from_numpy(out["energy_init_source"]).to(dev)
    c, dt = float(g["speed_of_sound"]), float(g["dt"])
    delay0 = torch.from_numpy(
        (out["distance_patches_to_source"] / c / dt).astype(np.int32)).to(dev)
    hist = exchange.energy_exchange(tables, e0, delay0, n_samples, int(g["max_order"]))
    etc = hist.dense().double().cpu().numpy()
    assert etc.shape == etc_ref.shape
    # per (patch-independent) band/direction tolerance on the whole histogram
    assert rel_err(etc, etc_ref) < TOL[dtype]
    # nothing may appear before the first possible arrival (delay bins bit-exact)
    assert np.array_equal(etc == 0, etc_ref == 0) or dtype == "f32"
    # golden (live reference) cross-check
    if "etc" in g:
        assert rel_err(etc, g["etc"]) < TOL[dtype]
    else:
        assert rel_err(etc[g["etc_rows"]], g["etc_sample"]) < TOL[dtype]

    # order 0 == initial energy only (RadiosityFast.py:550-555)
    h0 = exchange.energy_exchange(tables, e0, delay0, n_samples, 0)
    assert rel_err(h0.dense().double().cpu().numpy().sum(-1), g["etc_order0_sums"]) < TOL[dtype]

    # receiver collection
    air = g["air_attenuation"]
    rcv = g["receivers"]
    cen = out["patches_center"]
    dist = np.sqrt(((cen[None] - rcv[:, None]) ** 2).sum(-1))      # tolerance path
    scale = out["receiver_factor"][:, :, None] * np.exp(-air[None, None, :] * dist[:, :, None])
    shift = np.mod(out["receiver_delays"], n_samples).astype(np.int32)
    rdir = out["receiver_dir_index"].astype(np.int32)
    mono = exchange.collect_mono(
        hist, torch.from_numpy(rdir).to(dev), torch.from_numpy(shift).to(dev),
        torch.from_numpy(scale).to(dev))
    mono = mono.double().cpu().numpy()
    for r in range(mono.shape[0]):
        for b in range(mono.shape[1]):
            assert rel_err(mono[r, b], g["etc_receiver_mono"][r, b]) < TOL[dtype]
    pw = exchange.collect_patchwise(
        hist, torch.from_numpy(rdir).to(dev), torch.from_numpy(shift).to(dev),
        torch.from_numpy(scale).to(dev)).double().cpu().numpy()
    assert rel_err(pw.sum(-1), g["etc_receiver_patch_sums"]) < TOL[dtype]
    assert rel_err(pw.sum(1), mono) < 10 * TOL[dtype]


def test_exchange_linearity_and_empty(oracle):
    """Size-independent properties: linear in E0; no pairs -> only initial energy."""
    from sparrowpy_b200 import exchange
    g = load_golden("scene_occluder")
    out = oracle_run(oracle, g)
    dev = torch.device("cuda:0")
    n_samples = out["etc"].shape[-1]
    tables = device_tables(g, out, "f64", n_samples)
    e0 = torch.from_numpy(out["energy_init_source"]).to(dev)
    c, dt = float(g["speed_of_sound"]), float(g["dt"])
    delay0 = torch.from_numpy(
        (out["distance_patches_to_source"] / c / dt).astype(np.int32)).to(dev)
    a = exchange.energy_exchange(tables, e0, delay0, n_samples, 3).dense().clone()
    b = exchange.energy_exchange(tables, 2.5 * e0, delay0, n_samples, 3).dense().clone()
    assert torch.allclose(b, 2.5 * a, rtol=1e-12, atol=0)
    # empty pair list
    empty = exchange.build_pair_tables(
        torch.zeros(0, dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.int64, device=dev),
        torch.zeros(0, dtype=torch.float64, device=dev), torch.zeros(0, dtype=torch.int64, device=dev),
        torch.zeros(0, dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.int64, device=dev),
        torch.ones((1, 1, 1), dtype=torch.float64, device=dev), e0.shape[0], n_samples, "f64")
    h = exchange.energy_exchange(empty, e0, delay0, n_samples, 4).dense()
    h0 = exchange.energy_exchange(tables, e0, delay0, n_samples, 0).dense()
    assert torch.equal(h, h0)


def test_tiled_and_csr_gather_agree_on_ragged_lists():
    """Random sparse pair lists (empty segments, delays across bucket borders,
    N not a multiple of the receiver tile): both stage-1 kernels give the same G."""
    from sparrowpy_b200 import _lib, exchange
    dev = torch.device("cuda:0")
    gen = torch.Generator(device="cpu").manual_seed(5)
    n, d, b, c, t_len = 53, 3, 2, 4, 300
    m = 4000
    sender = torch.randint(0, n, (m,), generator=gen)
    receiver = torch.randint(0, n, (m,), generator=gen)
    key = sender * n + receiver
    _, first = np.unique(key.numpy(), return_index=True)      # one pair per (i, j)
    sel = torch.from_numpy(np.sort(first))
    sender, receiver = sender[sel].to(dev), receiver[sel].to(dev)
    m = sender.numel()
    ff = torch.rand(m, generator=gen, dtype=torch.float64).to(dev)
    delay = torch.randint(0, 140, (m,), generator=gen).to(dev)
    delay[::7] = 31 + (delay[::7] % 3)                        # straddle the bucket border
    out_dir = torch.randint(0, d, (m,), generator=gen).to(dev)
    cls = torch.randint(0, c, (m,), generator=gen).to(dev)
    cls[receiver % 5 == 0] = 1                                # leave some segments empty
    coef = torch.rand((c, d, b), generator=gen, dtype=torch.float64).to(dev)
    tables = exchange.build_pair_tables(sender, receiver, ff, delay, out_dir, cls, coef, n,
                                        t_len, "f64", gather="tma")
    t_pad, pad = _lib.exchange_layout(t_len, tables.max_delay, tables.dtype)
    ld = t_pad + pad
    prev = torch.zeros((b * n * d, ld), dtype=torch.float64, device=dev)
    prev[:, pad:pad + t_len] = torch.rand((b * n * d, t_len), generator=gen,
                                          dtype=torch.float64).to(dev)
    g1 = torch.zeros((b * c * n, ld), dtype=torch.float64, device=dev)
    g2 = torch.zeros_like(g1)
    st, code = _lib.stream_ptr(), _lib.I32(tables.dtype)
    _lib.call("spb_exchange_gather", prev, g1, tables.seg_ptr, tables.src, tables.wgt,
              tables.dly, n, n, c, d, b, 0, b, 0, n, t_pad, ld, pad, code, st)
    _lib.call("spb_exchange_gather_tiled", prev, g2, tables.ent_ptr, tables.recs, None, n, n,
              c, d, b, 0, b, 0, n, t_pad, ld, pad, code, st)
    # an explicit launch order (here: reversed) must not change anything
    n_tiles = c * (-(-n // 8))
    order = torch.arange(n_tiles - 1, -1, -1, dtype=torch.int32, device=dev)
    g3 = torch.zeros_like(g1)
    _lib.call("spb_exchange_gather_tiled", prev, g3, tables.ent_ptr, tables.recs, order, n, n,
              c, d, b, 0, b, 0, n, t_pad, ld, pad, code, st)
    torch.cuda.synchronize()
    assert torch.equal(g2, g3)
    torch.cuda.synchronize()
    a, bb = g1[:, pad:pad + t_len], g2[:, pad:pad + t_len]
    assert torch.allclose(a, bb, rtol=1e-12, atol=1e-14)
    assert a.abs().max() > 0
    # brute-force check of one non-empty segment
    seg = int(torch.nonzero(tables.seg_ptr[1:] - tables.seg_ptr[:-1])[0])
    cc, jj = seg // n, seg % n
    sel = (cls == cc) & (receiver == jj) & (delay < t_len)
    ref = torch.zeros(t_len, dtype=torch.float64, device=dev)
    for q in torch.nonzero(sel).reshape(-1).tolist():
        row = 1 * n * d + int(sender[q]) * d + int(out_dir[q])      # band 1
        dl = int(delay[q])
        ref[dl:] += ff[q] * prev[row, pad:pad + t_len - dl]
    assert torch.allclose(g2[1 * c * n + seg, pad:pad + t_len], ref, rtol=1e-12, atol=1e-14)


def test_sharded_driver_matches_one_call_api(oracle):
    """distributed.ShardedExchange (world = 1, per-order C-ABI calls) equals
    spb_energy_exchange (one call), and a 2-shard split run by hand equals both."""
    from sparrowpy_b200 import _lib, distributed, exchange
    g = load_golden("scene_directional")
    out = oracle_run(oracle, g)
    dev = torch.device("cuda:0")
    n_samples = out["etc"].shape[-1]
    tables = device_tables(g, out, "f64", n_samples)
    e0 = torch.from_numpy(out["energy_init_source"]).to(dev)
    c, dt = float(g["speed_of_sound"]), float(g["dt"])
    delay0 = torch.from_numpy(
        (out["distance_patches_to_source"] / c / dt).astype(np.int32)).to(dev)
    ref = exchange.energy_exchange(tables, e0, delay0, n_samples, 4).dense().clone()
    sx = distributed.ShardedExchange(tables, n_samples, dev)
    sx.init(e0, delay0)
    got = sx.run(4).dense()
    assert torch.equal(got, ref)
    # emulate two ranks on one GPU: each computes its receiver range, rows are merged
    n = tables.n_patches
    lo0, hi0, size = distributed.shard_range(n, 0, 2)
    lo1, hi1, _ = distributed.shard_range(n, 1, 2)
    assert (lo0, hi1) == (0, n) and hi0 == lo1
    sx.init(e0, delay0)
    prev, cur = sx.e_a, sx.e_b
    code, st = _lib.I32(tables.dtype), _lib.stream_ptr()
    t = tables
    for _ in range(4):
        for lo, hi in ((lo0, hi0), (lo1, hi1)):
            for b in range(t.n_bands):           # one band at a time, like the pipeline
                exchange.launch_gather(t, prev, sx.g, None, sx.n_alloc, b, b + 1, lo, hi,
                                       sx.t_pad, sx.ld, sx.pad)
                _lib.call("spb_exchange_mix", sx.g, cur, sx.e_total, t.seg_ptr, t.coef,
                          t.n_patches, sx.n_alloc, t.n_classes, t.n_dirs, t.n_bands, b, b + 1,
                          lo, hi, sx.t_pad, sx.ld, sx.pad, code, st)
        prev, cur = cur, prev
    got2 = exchange.EnergyHistogram(sx.e_total, n, t.n_dirs, t.n_bands, n_samples, sx.pad,
                                    n_alloc=sx.n_alloc).dense()
    assert torch.equal(got2, ref)


def test_exchange_invariant_under_patch_renumbering(oracle):
    """Any internal patch numbering (with holes) gives the same histogram and the
    same receiver ETC in the caller's numbering."""
    from sparrowpy_b200 import exchange
    g = load_golden("scene_directional")
    out = oracle_run(oracle, g)
    dev = torch.device("cuda:0")
    n_samples = out["etc"].shape[-1]
    base = device_tables(g, out, "f64", n_samples)
    n = base.n_patches
    e0 = torch.from_numpy(out["energy_init_source"]).to(dev)
    c, dt = float(g["speed_of_sound"]), float(g["dt"])
    delay0 = torch.from_numpy(
        (out["distance_patches_to_source"] / c / dt).astype(np.int32)).to(dev)
    ref = exchange.energy_exchange(base, e0, delay0, n_samples, 4)
    gen = np.random.default_rng(11)
    n_int = n + 37
    rank = torch.from_numpy(np.sort(gen.choice(n_int, n, replace=False))[gen.permutation(n)])

    # rebuild the tables with the renumbering (same inputs as device_tables)
    import sparrowpy_b200.exchange as ex
    pairs = torch.from_numpy(out["visible_patches"]).to(dev)
    ff = torch.from_numpy(out["ff_pairs"]).to(dev)
    areas = torch.from_numpy(out["patches_area"]).to(dev)
    sender, receiver, ffd = ex.directed_pairs(pairs, ff, areas)
    s_in = g["vi"].shape[1]
    brdf = g["brdf"].reshape(g["brdf"].shape[0], s_in, g["vo"].shape[1], -1)
    wall = torch.from_numpy(out["patch_to_wall_ids"]).to(dev)
    cls = torch.from_numpy(g["brdf_index"]).to(dev)[wall[sender]] * s_in + \
        torch.from_numpy(out["in_dir"]).to(dev)
    coef = torch.from_numpy(np.exp(-g["air_attenuation"])[None, None, :] * brdf.reshape(
        -1, brdf.shape[2], brdf.shape[3])).to(dev)
    tables = ex.build_pair_tables(
        sender, receiver, ffd, torch.from_numpy(out["pair_delays"]).to(dev),
        torch.from_numpy(out["out_dir"]).to(dev), cls, coef, n, n_samples, "f64",
        rank=rank.to(dev), n_internal=n_int)
    assert tables.n_patches == n_int and tables.n_user == n
    got = exchange.energy_exchange(tables, e0, delay0, n_samples, 4)
    a, b = got.dense(), ref.dense()
    assert a.shape == b.shape
    assert float((a - b).abs().max() / b.abs().max()) < 1e-13
    # receiver collection takes per-patch inputs in the caller's numbering
    air, rcv, cen = g["air_attenuation"], g["receivers"], out["patches_center"]
    dist = np.sqrt(((cen[None] - rcv[:, None]) ** 2).sum(-1))
    scale = torch.from_numpy(out["receiver_factor"][:, :, None] *
                             np.exp(-air[None, None, :] * dist[:, :, None])).to(dev)
    shift = torch.from_numpy(np.mod(out["receiver_delays"], n_samples).astype(np.int32)).to(dev)
    rdir = torch.from_numpy(out["receiver_dir_index"].astype(np.int32)).to(dev)
    m1 = exchange.collect_mono(got, rdir, shift, scale)
    m0 = exchange.collect_mono(ref, rdir, shift, scale)
    assert float((m1 - m0).abs().max() / m0.abs().max()) < 1e-12
    p1 = exchange.collect_patchwise(got, rdir, shift, scale)
    p0 = exchange.collect_patchwise(ref, rdir, shift, scale)
    assert p1.shape == p0.shape
    assert float((p1 - p0).abs().max() / p0.abs().max()) < 1e-12


@pytest.mark.parametrize("name", ["scene_c1", "scene_canyon01", "scene_occluder"])
def test_fused_order_equals_gather_plus_mix(oracle, name, monkeypatch):
    """Diffuse scenes (one BRDF class, one direction): the whole reflection order in one
    kernel (spb_exchange_order_fused: stage 2 in the tensor-memory gather's epilogue) is
    bit-identical to gather + mix, including rows of receivers without pairs and a receiver
    sub-range (the sharded case)."""
    from sparrowpy_b200 import distributed, exchange
    monkeypatch.setenv("SPB_GATHER", "tmem")
    g = load_golden(name)
    out = oracle_run(oracle, g)
    n_samples = out["etc"].shape[-1]
    dev = torch.device("cuda:0")
    tables = device_tables(g, out, "f64", n_samples)
    assert tables.n_classes == 1 and tables.n_dirs == 1 and tables.win_recs is not None
    e0 = torch.from_numpy(out["energy_init_source"]).to(dev)
    c, dt = float(g["speed_of_sound"]), float(g["dt"])
    delay0 = torch.from_numpy(
        (out["distance_patches_to_source"] / c / dt).astype(np.int32)).to(dev)
    orders = int(g["max_order"])
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("SPB_FUSED_ORDER", flag)
        sx = distributed.ShardedExchange(tables, n_samples, dev, local=True)
        assert sx.fused_order() == (flag == "1")
        # stale data in the ping-pong buffers must not survive (receivers without pairs)
        sx.init(e0, delay0)
        sx.e_b[:, sx.pad:] = 7.0
        res[flag] = sx.run(orders).dense().clone()
    assert torch.equal(res["1"], res["0"])
    assert rel_err(res["1"].cpu().numpy(), out["etc"]) < 1e-6
