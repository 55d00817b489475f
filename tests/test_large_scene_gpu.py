"""Band-by-band, sharded-total schedule (distributed.BandwiseExchange) on one GPU against
the ordinary exchange, on the two-band golden scene.  The schedule's choreography is
covered on the CPU (tests/test_distributed_cpu.py); this runs it with the CUDA kernels."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from test_exchange_gpu import device_tables, oracle_run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("band_block", [1, 2])
def test_bandwise_equals_ordinary_exchange(oracle, band_block):
    from sparrowpy_b200 import distributed, exchange
    g = load_golden("scene_directional")
    out = oracle_run(oracle, g)
    n_samples = out["etc"].shape[-1]
    dev = torch.device("cuda:0")
    tables = device_tables(g, out, "f64", n_samples)
    e0 = torch.from_numpy(out["energy_init_source"]).to(dev)
    c, dt = float(g["speed_of_sound"]), float(g["dt"])
    delay0 = torch.from_numpy(
        (out["distance_patches_to_source"] / c / dt).astype(np.int32)).to(dev)
    orders = int(g["max_order"])
    ref_hist = exchange.energy_exchange(tables, e0, delay0, n_samples, orders)
    ref = ref_hist.dense().clone()
    bx = distributed.BandwiseExchange(tables, n_samples, dev, band_block=band_block)
    hist = bx.run(e0, delay0, orders)
    assert torch.equal(hist.dense_local(), ref)
    # sharded collection == ordinary collection
    air, rcv, cen = g["air_attenuation"], g["receivers"], out["patches_center"]
    dist = np.sqrt(((cen[None] - rcv[:, None]) ** 2).sum(-1))
    scale = torch.from_numpy(out["receiver_factor"][:, :, None]
                             * np.exp(-air[None, None, :] * dist[:, :, None])).to(dev)
    shift = torch.from_numpy(np.mod(out["receiver_delays"], n_samples).astype(np.int32)).to(dev)
    rdir = torch.from_numpy(out["receiver_dir_index"].astype(np.int32)).to(dev)
    want = exchange.collect_mono(ref_hist, rdir, shift, scale)
    got = hist.collect_mono(rdir, shift, scale)
    assert torch.allclose(got, want, rtol=1e-12, atol=0)
