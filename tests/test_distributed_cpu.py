"""Receiver-sharded exchange choreography (distributed.ShardedExchange) under gloo
with world_size 2 on CPU: shard ranges, padded all-gather per order, final gather.
The per-order local step is a CPU stand-in with the same contract as the CUDA
kernels (gather + mix restricted to [j_lo, j_hi))."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def make_tables(seed=3, n=37, d=2, b=2, c=3, t_len=90):
    from sparrowpy_b200 import exchange
    gen = torch.Generator().manual_seed(seed)
    m = 900
    sender = torch.randint(0, n, (m,), generator=gen)
    receiver = torch.randint(0, n, (m,), generator=gen)
    _, first = np.unique((sender * n + receiver).numpy(), return_index=True)
    sel = torch.from_numpy(np.sort(first))
    sender, receiver = sender[sel], receiver[sel]
    m = sender.numel()
    ff = torch.rand(m, generator=gen, dtype=torch.float64) * 0.05
    delay = torch.randint(0, 40, (m,), generator=gen)
    out_dir = torch.randint(0, d, (m,), generator=gen)
    cls = torch.randint(0, c, (m,), generator=gen)
    coef = torch.rand((c, d, b), generator=gen, dtype=torch.float64)
    tables = exchange.build_pair_tables(sender, receiver, ff, delay, out_dir, cls, coef, n,
                                        t_len, "f64")
    e0 = torch.rand((n, d, b), generator=gen, dtype=torch.float64)
    delay0 = torch.randint(0, 30, (n,), generator=gen).to(torch.int32)
    return tables, e0, delay0, t_len


def cpu_order(sx):
    """CPU stand-in for spb_exchange_gather + spb_exchange_mix on [j_lo, j_hi)."""
    def compute(prev, cur, total, b_lo, b_hi):
        t = sx.t            # looked up per call: the band-block schedule swaps the tables
        n, nd, nc = t.n_patches, t.n_dirs, t.n_classes
        seg_ptr = t.seg_ptr.tolist()
        for j in range(sx.j_lo, sx.j_hi):
            for b in range(b_lo, b_hi):
                band0 = b * sx.n_alloc * nd
                acc = torch.zeros((nd, sx.t_pad), dtype=prev.dtype)
                for c in range(nc):
                    seg = c * n + j
                    g = torch.zeros(sx.t_pad, dtype=prev.dtype)
                    for q in range(seg_ptr[seg], seg_ptr[seg + 1]):
                        row = band0 + int(t.src[q])
                        dl = int(t.dly[q])
                        g += t.wgt[q] * prev[row, sx.pad - dl: sx.pad - dl + sx.t_pad]
                    acc += t.coef[c, :, b][:, None] * g[None, :]
                for dd in range(nd):
                    row = band0 + j * nd + dd
                    cur[row, sx.pad:] = acc[dd]
                    total[row, sx.pad:] += acc[dd]
    return compute


def _worker(rank, world, port, orders, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparrowpy_b200 import distributed
    tables, e0, delay0, t_len = make_tables()
    sx = distributed.ShardedExchange(tables, t_len, torch.device("cpu"))
    sx.compute = cpu_order(sx)
    sx.init(e0, delay0)
    hist = sx.run(orders)
    if rank == 0:
        torch.save(hist.dense().clone(), out_path)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_ranges_cover_all_receivers():
    from sparrowpy_b200.distributed import shard_range
    for n in (1, 7, 8, 37, 148, 3700, 19200, 100000):
        for world in (1, 2, 4, 8):
            covered = []
            for r in range(world):
                lo, hi, size = shard_range(n, r, world)
                assert (lo % 8 == 0 or lo == hi) and size % 8 == 0 and 0 <= lo <= hi <= n
                covered += list(range(lo, hi))
            assert covered == list(range(n))


@pytest.mark.timeout(300)
def test_sharded_exchange_world2_equals_single(tmp_path):
    orders = 3
    # single process (no process group): world = 1
    from sparrowpy_b200 import distributed
    tables, e0, delay0, t_len = make_tables()
    sx = distributed.ShardedExchange(tables, t_len, torch.device("cpu"))
    sx.compute = cpu_order(sx)
    sx.init(e0, delay0)
    single = sx.run(orders).dense().clone()
    assert single.abs().sum() > 0
    out = str(tmp_path / "world2.pt")
    mp.spawn(_worker, args=(2, _free_port(), orders, out), nprocs=2, join=True)
    sharded = torch.load(out)
    assert torch.equal(single, sharded)
    # order 0 places e0 at the source delay bins
    sx.init(e0, delay0)
    h0 = sx.run(0).dense()
    for i in (0, 5, 36):
        assert torch.equal(h0[i, :, :, int(delay0[i])], e0[i])


# ---------------------------------------------------------------------------
# large-scene schedule: band blocks + sharded E_total + sharded receiver collection
# ---------------------------------------------------------------------------
def cpu_collect(hist, rdir, shift, scale):
    """CPU stand-in for exchange.collect_mono (circular roll, RadiosityFast.py:1181-1183)."""
    dense = hist.dense()                                   # (n, D, B, T)
    n_rcv, n = rdir.shape
    out = torch.zeros((n_rcv, hist.n_bands, hist.n_samples), dtype=dense.dtype)
    for r in range(n_rcv):
        for k in range(n):
            for b in range(hist.n_bands):
                sc = scale[r, k, b]
                if sc != 0:
                    out[r, b] += sc * torch.roll(dense[k, int(rdir[r, k]), b], int(shift[r, k]))
    return out


def receiver_inputs(n, d, b, t_len, seed=11):
    gen = torch.Generator().manual_seed(seed)
    rdir = torch.randint(0, d, (2, n), generator=gen).to(torch.int32)
    shift = torch.randint(0, t_len, (2, n), generator=gen).to(torch.int32)
    scale = torch.rand((2, n, b), generator=gen, dtype=torch.float64)
    scale[:, ::3] = 0                                       # invisible patches
    return rdir, shift, scale


def _bandwise_worker(rank, world, port, orders, band_block, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparrowpy_b200 import distributed
    tables, e0, delay0, t_len = make_tables(b=4)
    bx = distributed.BandwiseExchange(tables, t_len, torch.device("cpu"),
                                      band_block=band_block, collect=cpu_collect)
    bx.sx.compute = cpu_order(bx.sx)
    hist = bx.run(e0, delay0, orders)
    mono = hist.collect_mono(*receiver_inputs(tables.n_patches, tables.n_dirs,
                                              tables.n_bands, t_len))
    torch.save({"local": hist.dense_local().clone(), "j": (hist.j_lo, hist.j_hi),
                "mono": mono}, f"{out_path}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("band_block", [1, 2])
def test_bandwise_sharded_schedule_equals_single(tmp_path, band_block):
    from sparrowpy_b200 import distributed
    orders = 2
    tables, e0, delay0, t_len = make_tables(b=4)
    sx = distributed.ShardedExchange(tables, t_len, torch.device("cpu"))
    sx.compute = cpu_order(sx)
    sx.init(e0, delay0)
    full_hist = sx.run(orders)
    full = full_hist.dense().clone()
    rcv = receiver_inputs(tables.n_patches, tables.n_dirs, tables.n_bands, t_len)
    mono_ref = cpu_collect(full_hist, *rcv)
    # one process: band blocks only
    bx = distributed.BandwiseExchange(tables, t_len, torch.device("cpu"),
                                      band_block=band_block, collect=cpu_collect)
    bx.sx.compute = cpu_order(bx.sx)
    h1 = bx.run(e0, delay0, orders)
    assert torch.equal(h1.dense_local(), full)
    assert torch.allclose(h1.collect_mono(*rcv), mono_ref, rtol=1e-13, atol=0)
    # two ranks: every rank keeps its own receivers only; the collection is all-reduced
    out = str(tmp_path / "bw")
    mp.spawn(_bandwise_worker, args=(2, _free_port(), orders, band_block, out), nprocs=2,
             join=True)
    parts = [torch.load(f"{out}.{r}") for r in range(2)]
    got = torch.cat([p["local"] for p in parts])
    assert [p["j"] for p in parts] == [(0, 24), (24, 37)]
    assert torch.equal(got, full)
    for p in parts:
        assert torch.allclose(p["mono"], mono_ref, rtol=1e-13, atol=0)


def test_shard_restricted_tables_partition_the_full_tables():
    """build_pair_tables(receiver_range=...) keeps exactly the rank's segments; layout
    quantities stay global."""
    from sparrowpy_b200 import exchange
    from sparrowpy_b200.distributed import shard_range
    from test_tables_cpu import random_pairs
    sender, receiver, ff, delay, out_dir, cls, coef, n, t_len = random_pairs(7)
    full = exchange.build_pair_tables(sender, receiver, ff, delay, out_dir, cls, coef, n, t_len,
                                      "f64")
    world = 3
    seen = 0
    for r in range(world):
        lo, hi, _ = shard_range(n, r, world)
        part = exchange.build_pair_tables(sender, receiver, ff, delay, out_dir, cls, coef, n,
                                          t_len, "f64", receiver_range=(lo, hi))
        assert (part.max_delay, part.n_directed) == (full.max_delay, full.n_directed)
        cnt_full = (full.seg_ptr[1:] - full.seg_ptr[:-1]).view(full.n_classes, n)
        cnt_part = (part.seg_ptr[1:] - part.seg_ptr[:-1]).view(full.n_classes, n)
        assert torch.equal(cnt_part[:, lo:hi], cnt_full[:, lo:hi])
        assert int(cnt_part.sum()) == int(cnt_full[:, lo:hi].sum())
        for c in range(full.n_classes):
            for j in range(lo, hi):
                a0, a1 = full.seg_ptr[c * n + j], full.seg_ptr[c * n + j + 1]
                b0, b1 = part.seg_ptr[c * n + j], part.seg_ptr[c * n + j + 1]
                assert torch.equal(full.src[a0:a1], part.src[b0:b1])
                assert torch.equal(full.dly[a0:a1], part.dly[b0:b1])
                assert torch.equal(full.wgt[a0:a1], part.wgt[b0:b1])
        # tile records of the shard's tiles are those of the full tables
        rec_full = (full.tile_ptr[1:] - full.tile_ptr[:-1]).view(full.n_classes, -1)
        rec_part = (part.tile_ptr[1:] - part.tile_ptr[:-1]).view(full.n_classes, -1)
        assert torch.equal(rec_part[:, lo // 8:-(-hi // 8)], rec_full[:, lo // 8:-(-hi // 8)])
        seen += int(cnt_part.sum())
    assert seen == full.src.numel()


def _route_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparrowpy_b200 import distributed
    gen = torch.Generator().manual_seed(100 + rank)
    m = 50 + 30 * rank                                  # ragged: ranks post different counts
    ints = torch.randint(0, 1000, (m, 5), generator=gen, dtype=torch.int32)
    ints[:, 0] = rank                                   # who posted it
    ff = torch.rand(m, generator=gen, dtype=torch.float64)
    dest = torch.randint(0, world, (m,), generator=gen)
    got_i, got_f = distributed.route_directed(dest, ints, ff, world)
    torch.save((dest, ints, ff, got_i, got_f), os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_route_directed_delivers_every_pair_to_its_owner(tmp_path):
    """The all-to-all of the sharded bake (distributed.route_directed) under gloo, world 3:
    every posted (int record, weight) arrives exactly once at the rank it was addressed to,
    records and weights stay paired."""
    world = 3
    mp.spawn(_route_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    data = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    for r in range(world):
        want = []
        for dest, ints, ff, _, _ in data:
            sel = dest == r
            want += [(tuple(i.tolist()), float(f)) for i, f in zip(ints[sel], ff[sel])]
        got = [(tuple(i.tolist()), float(f)) for i, f in zip(data[r][3], data[r][4])]
        assert sorted(got) == sorted(want) and len(got) > 0
    # world 1: a pure reordering by destination
    from sparrowpy_b200 import distributed
    ints = torch.arange(20, dtype=torch.int32).reshape(4, 5)
    gi, gf = distributed.route_directed(torch.zeros(4, dtype=torch.long), ints,
                                        torch.arange(4, dtype=torch.float64), 1)
    assert torch.equal(gi, ints) and gf.tolist() == [0.0, 1.0, 2.0, 3.0]
