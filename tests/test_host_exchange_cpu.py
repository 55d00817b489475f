"""The host side of the exchange end to end on the CPU: pair tables built from the
oracle's bake (directed pairs, reverse form factors, BRDF classes, compact patch
renumbering with holes, tile records) driven through distributed.ShardedExchange with a
CPU stand-in for the two kernels reproduce the oracle's (and the live reference's) ETC.
This pins everything between the bake kernels and the exchange kernels -- the table
semantics the CUDA kernels are written against -- without a GPU."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from test_distributed_cpu import cpu_order
from test_exchange_gpu import oracle_run


def cpu_tables(g, out, n_samples, renumber):
    from sparrowpy_b200 import exchange, geometry
    n = out["patches_center"].shape[0]
    pairs = torch.from_numpy(out["visible_patches"])
    ff = torch.from_numpy(out["ff_pairs"])
    areas = torch.from_numpy(out["patches_area"])
    sender, receiver, ffd = exchange.directed_pairs(pairs, ff, areas)
    delay = torch.from_numpy(out["pair_delays"])
    out_dir = torch.from_numpy(out["out_dir"])
    if "brdf_dirs" in g:
        s_in = g["vi"].shape[1]
        brdf = g["brdf"].reshape(g["brdf"].shape[0], s_in, g["vo"].shape[1], -1)
        wall = torch.from_numpy(out["patch_to_wall_ids"])
        bidx = torch.from_numpy(g["brdf_index"])
        cls = bidx[wall[sender]] * s_in + torch.from_numpy(out["in_dir"])
        coef = np.exp(-g["air_attenuation"])[None, None, :] * brdf.reshape(
            -1, brdf.shape[2], brdf.shape[3])
    else:
        cls = torch.zeros_like(sender)
        coef = np.ones((1, 1, 1))
    coef = torch.from_numpy(np.ascontiguousarray(coef))
    rank = n_int = None
    if renumber:
        r, n_int = geometry.compact_patch_order(g["patches_points"], g["patch_to_wall_ids"])
        rank = torch.from_numpy(r)
    return exchange.build_pair_tables(sender, receiver, ffd, delay, out_dir, cls, coef, n,
                                      n_samples, "f64", rank=rank, n_internal=n_int)


@pytest.mark.parametrize("renumber", [False, True])
@pytest.mark.parametrize("name", ["scene_uneven", "scene_cube05"])
def test_tables_and_driver_reproduce_the_oracle_etc(oracle, name, renumber):
    from sparrowpy_b200 import distributed
    g = load_golden(name)
    out = oracle_run(oracle, g)
    etc_ref = out["etc"]
    n_samples = etc_ref.shape[-1]
    tables = cpu_tables(g, out, n_samples, renumber)
    if renumber:
        assert tables.n_patches >= tables.n_user and tables.n_patches % 8 == 0
    e0 = torch.from_numpy(out["energy_init_source"])
    c, dt = float(g["speed_of_sound"]), float(g["dt"])
    delay0 = torch.from_numpy((out["distance_patches_to_source"] / c / dt).astype(np.int32))
    sx = distributed.ShardedExchange(tables, n_samples, torch.device("cpu"))
    sx.compute = cpu_order(sx)
    sx.init(e0, delay0)
    etc = sx.run(int(g["max_order"])).dense().numpy()
    assert etc.shape == etc_ref.shape
    assert rel_err(etc, etc_ref) < 1e-12
    assert np.array_equal(etc == 0, etc_ref == 0)
    if "etc" in g:                       # the live reference's histogram
        assert rel_err(etc, g["etc"]) < 1e-12
