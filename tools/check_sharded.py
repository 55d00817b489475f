"""Multi-GPU parity check (run under torchrun): the receiver-sharded exchange in
every communication mode equals the single-GPU exchange computed on the same rank.

    python -m torch.distributed.run --nproc-per-node 2 tools/check_sharded.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sparrowpy_b200 import bake, distributed, exchange  # noqa: E402

rank = int(os.environ["RANK"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = bench.CONFIGS[os.environ.get("CHECK_CONFIG", "c2s")]
orders = 4
ok = True
for dtype in ("f64", "f32"):
    rad = bench.build_scene(cfg, dtype)
    n_samples = cfg["n_samples"]
    world = dist.get_world_size()
    delay0 = bake.delay_bins(rad._d0_dev, bench.SPEED_OF_SOUND, bench.DT)
    # single-GPU reference with the default patch numbering ...
    tables1 = rad._pair_tables(bench.SPEED_OF_SOUND, bench.DT, n_samples)
    ref1 = exchange.energy_exchange(tables1, rad._e0_dev, delay0, n_samples, orders).dense().clone()
    # ... and with the shard-balanced numbering the multi-GPU run uses (same tiles, so
    # the sharded result must be bit-identical to this one)
    tables = rad._pair_tables(bench.SPEED_OF_SOUND, bench.DT, n_samples, n_shards=world)
    ref = exchange.energy_exchange(tables, rad._e0_dev, delay0, n_samples, orders).dense().clone()
    renum = float((ref - ref1).abs().max() / ref1.abs().max())
    ok &= renum < (1e-12 if dtype == "f64" else 1e-5)
    if rank == 0:
        print(f"{dtype} renumbering for {world} shards: max rel diff {renum:.2e}", flush=True)
    for mode in ("multicast", "p2p", "nccl"):
        os.environ["SPB_COMM"] = mode
        sx = distributed.ShardedExchange(tables, n_samples, dev)
        for rep in range(2):                      # twice: buffer reuse across steps
            sx.init(rad._e0_dev, delay0)
            got = sx.run(orders).dense()
            torch.cuda.synchronize()
            same = bool(torch.equal(got, ref))
            err = float((got - ref).abs().max() / ref.abs().max())
            ok &= same
            if rank == 0:
                print(f"{dtype} {mode:9s} (ran as {sx.comm:9s}) rep {rep}: "
                      f"equal={same} max rel diff {err:.2e}", flush=True)
        del sx
    # large-scene schedule: shard-restricted tables, band by band, E_total sharded
    os.environ["SPB_COMM"] = "p2p"
    part = rad._pair_tables(bench.SPEED_OF_SOUND, bench.DT, n_samples, n_shards=world,
                            shard=rank)
    # the sharded bake (rows of the visibility matrix per rank + all-to-all of the directed
    # pairs) must build exactly these tables
    sb, n_pairs = distributed.sharded_bake_tables(rad, bench.SPEED_OF_SOUND, bench.DT, n_samples)
    same = n_pairs == rad._baked["pairs"].shape[0] and all(
        torch.equal(getattr(sb, f), getattr(part, f))
        for f in ("seg_ptr", "src", "wgt", "dly", "coef", "rank") + (
            ("win_ptr", "win_recs") if part.win_recs is not None else ("ent_ptr", "recs")))
    ok &= bool(same)
    if rank == 0:
        print(f"{dtype} sharded bake: tables equal={bool(same)}", flush=True)
    del sb
    for block in (1, 2):
        bx = distributed.BandwiseExchange(part, n_samples, dev, band_block=block)
        h = bx.run(rad._e0_dev, delay0, orders)
        torch.cuda.synchronize()
        inv = torch.full((tables.n_patches,), -1, dtype=torch.long, device=dev)
        inv[tables.rank] = torch.arange(tables.n_user, device=dev)
        own = inv[h.j_lo:h.j_hi]
        want = torch.zeros_like(h.dense_local())
        want[own >= 0] = ref[own[own >= 0]]
        same = bool(torch.equal(h.dense_local(), want))
        ok &= same
        if rank == 0:
            print(f"{dtype} bandwise block={block}: own rows equal={same}", flush=True)
        del bx, h
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED PARITY", "OK" if int(flag.item()) == 1 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
