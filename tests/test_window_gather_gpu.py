"""Tensor-memory gather (csrc/exchange_tmem.cu, the default FP64 kernel) vs the CSR kernel
and the oracle: ragged random pair lists for both window widths, several histogram
lengths and band counts (every mapping of the TMEM lane quarters), receiver/band ranges,
launch orders; then golden scenes end to end."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from test_exchange_gpu import device_tables, oracle_run

pytestmark = pytest.mark.gpu



def ragged(seed, n, d, c, m, max_delay, spread, dev):
    """Random pairs whose delays are clustered per (sender, receiver tile) so that the
    records really hold several receivers."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sender = torch.randint(0, n, (m,), generator=gen)
    receiver = torch.randint(0, n, (m,), generator=gen)
    _, first = np.unique((sender * n + receiver).numpy(), return_index=True)
    sel = torch.from_numpy(np.sort(first))
    sender, receiver = sender[sel], receiver[sel]
    m = sender.numel()
    centre = torch.randint(0, max_delay - spread, (n, (n + 7) // 8), generator=gen)
    delay = centre[sender, receiver // 8] + torch.randint(0, spread + 1, (m,), generator=gen)
    ff = torch.rand(m, generator=gen, dtype=torch.float64)
    out_dir = torch.randint(0, d, (m,), generator=gen)
    cls = torch.randint(0, c, (m,), generator=gen)
    cls[receiver % 5 == 0] = 1                                # leave some segments empty
    return [x.to(dev) for x in (sender, receiver, ff, delay, out_dir, cls)], gen


@pytest.mark.parametrize("t_len,spread,width", [
    (300, 3, 4), (300, 9, 10), (300, 25, 10), (1000, 9, 4), (1200, 6, 10), (2000, 9, 10),
    (2300, 10, 10), (5000, 4, 4)])
@pytest.mark.parametrize("n_bands", [1, 2, 5])
def test_tmem_gather_equals_csr_gather(t_len, spread, width, n_bands, monkeypatch):
    """Tensor-memory gather (csrc/exchange_tmem.cu) vs the CSR kernel: every mapping of the
    four TMEM lane quarters onto (band, 512-bin chunk) units -- 4 bands x 1 chunk, 2 x 2,
    1 x 4, several chunks per band, ragged last chunk -- and receiver / band ranges."""
    from sparrowpy_b200 import _lib, exchange
    monkeypatch.setenv("SPB_GATHER", "tmem")
    dev = torch.device("cuda:0")
    n, d, b, c = 53, 3, n_bands, 4
    (sender, receiver, ff, delay, out_dir, cls), gen = ragged(
        t_len + spread, n, d, c, 2500, min(t_len, 400) + 20, spread, dev)
    coef = torch.rand((c, d, b), generator=gen, dtype=torch.float64).to(dev)
    tables = exchange.build_pair_tables(sender, receiver, ff, delay, out_dir, cls, coef, n,
                                        t_len, "f64")
    assert tables.win_recs is not None and tables.recs is None
    keep = delay < t_len
    tables.win_ptr, tables.win_recs, tables.win_w = exchange.build_window_records(
        *(x[keep] for x in (sender, receiver, ff, delay, out_dir, cls)), n, d, c, _lib.F64,
        width=width)
    tables.win_ptr, tables.win_recs = exchange.device_window_records(
        tables.win_ptr, tables.win_recs)
    t_pad, pad = _lib.exchange_layout(t_len, tables.max_delay, tables.dtype)
    ld = t_pad + pad
    n_alloc = n + 3                                  # padded patch axis (multi-GPU layout)
    prev = torch.zeros((b * n_alloc * d, ld), dtype=torch.float64, device=dev)
    prev[:, pad:pad + t_len] = torch.rand((b * n_alloc * d, t_len), generator=gen,
                                          dtype=torch.float64).to(dev)
    g1 = torch.zeros((b * c * n, ld), dtype=torch.float64, device=dev)
    g2 = torch.zeros_like(g1)
    st, code = _lib.stream_ptr(), _lib.I32(tables.dtype)
    _lib.call("spb_exchange_gather", prev, g1, tables.seg_ptr, tables.src, tables.wgt,
              tables.dly, n, n_alloc, c, d, b, 0, b, 0, n, t_pad, ld, pad, code, st)
    exchange.launch_gather(tables, prev, g2, None, n_alloc, 0, b, 0, n, t_pad, ld, pad)
    torch.cuda.synchronize()
    a, bb = g1[:, pad:pad + t_len], g2[:, pad:pad + t_len]
    assert a.abs().max() > 0
    assert torch.allclose(a, bb, rtol=1e-12, atol=1e-13)
    n_tiles = c * (-(-n // 8))
    order = torch.arange(n_tiles - 1, -1, -1, dtype=torch.int32, device=dev)
    g3 = torch.zeros_like(g1)
    exchange.launch_gather(tables, prev, g3, order, n_alloc, 0, b, 0, n, t_pad, ld, pad)
    torch.cuda.synchronize()
    assert torch.equal(g2, g3)
    if b >= 2:
        g4 = torch.zeros_like(g1)
        exchange.launch_gather(tables, prev, g4, None, n_alloc, 1, b, 16, 40, t_pad, ld, pad)
        torch.cuda.synchronize()
        g4v, g2v = g4.view(b, c, n, ld), g2.view(b, c, n, ld)
        assert torch.equal(g4v[1:, :, 16:40], g2v[1:, :, 16:40])
        assert float(g4v[0].abs().max()) == 0 and float(g4v[1:, :, :16].abs().max()) == 0
        assert float(g4v[1:, :, 40:].abs().max()) == 0


@pytest.mark.parametrize("name", ["scene_c1", "scene_directional", "scene_canyon01"])
def test_tmem_exchange_matches_oracle(oracle, name, monkeypatch):
    from sparrowpy_b200 import exchange
    monkeypatch.setenv("SPB_GATHER", "tmem")
    g = load_golden(name)
    out = oracle_run(oracle, g)
    etc_ref = out["etc"]
    n_samples = etc_ref.shape[-1]
    dev = torch.device("cuda:0")
    tables = device_tables(g, out, "f64", n_samples)
    assert tables.win_recs is not None
    e0 = torch.from_numpy(out["energy_init_source"]).to(dev)
    c, dt = float(g["speed_of_sound"]), float(g["dt"])
    delay0 = torch.from_numpy(
        (out["distance_patches_to_source"] / c / dt).astype(np.int32)).to(dev)
    hist = exchange.energy_exchange(tables, e0, delay0, n_samples, int(g["max_order"]))
    etc = hist.dense().double().cpu().numpy()
    assert rel_err(etc, etc_ref) < 1e-6
    assert np.array_equal(etc == 0, etc_ref == 0)
