#!/usr/bin/env python
"""One-call comparison of the stage-1 gather kernels on a bench scene (round-2 tool):
bakes once, then for every kernel variant builds its tables, runs a few reflection
orders, times the gather launches with CUDA events and checks the histogram against
the default kernel.

    python tools/sweep_gather.py --config c4 [--orders 3] [--variants tma,tmem,csr]"""
import argparse
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

ENV = {
    "tma": dict(SPB_GATHER="tma"),
    "tmem": dict(SPB_GATHER="tmem"),
    "csr": dict(SPB_GATHER="csr"),
}
KEYS = ("SPB_GATHER",)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--orders", type=int, default=3)
    ap.add_argument("--variants", default="tma,tmem")
    args = ap.parse_args()
    import torch
    import bench
    from sparrowpy_b200 import _lib, bake, distributed, exchange
    cfg = bench.CONFIGS[args.config]
    rad = bench.build_scene(cfg, "f64")
    dev = torch.device("cuda", 0)
    n_samples = cfg["n_samples"]
    e0 = rad._e0_dev.double().contiguous()
    delay0 = bake.delay_bins(rad._d0_dev, bench.SPEED_OF_SOUND, bench.DT)
    ref = None
    print(f"{args.config}: N={rad.n_patches} pairs={rad._baked['pairs'].shape[0]}", flush=True)
    for name in args.variants.split(","):
        base = name
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(ENV[base])
        rad._tables = None                         # tables depend on the kernel
        t0 = time.time()
        tables = rad._pair_tables(bench.SPEED_OF_SOUND, bench.DT, n_samples)
        torch.cuda.synchronize()
        t_tab = time.time() - t0
        sx = distributed.ShardedExchange(tables, n_samples, dev)
        events = []

        def gather_only(prev, cur, total, b_lo, b_hi, sx=sx, tables=tables):
            t = tables
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            exchange.launch_gather(t, prev, sx.g, sx.cta_order(), sx.n_alloc, b_lo, b_hi,
                                   sx.j_lo, sx.j_hi, sx.t_pad, sx.ld, sx.pad)
            ev[1].record()
            events.append(ev)
            sx._mix(cur, total, b_lo, b_hi)

        sx.compute = gather_only
        sx.init(e0, delay0)
        out = sx.run(args.orders).dense()
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in events]
        if ref is None:
            ref, err = out.clone(), 0.0
        else:
            err = float((out - ref).abs().max() / ref.abs().max())
        print(f"{name:12s} records={tables.n_records:9d} window={tables.win_w:2d} "
              f"tables {t_tab:5.1f}s  gather ms/launch: "
              + " ".join(f"{x:7.2f}" for x in ms) + f"   max rel diff vs first {err:.1e}",
              flush=True)
        del sx, tables, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
