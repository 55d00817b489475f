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
    t = sx.t

    def compute(prev, cur, total, b_lo, b_hi):
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
