#!/usr/bin/env python
"""One-call comparison of the stage-1 gather kernels on a bench scene (round-2 tool):
bakes once, then for every kernel variant builds its tables, runs a few reflection
orders, times the gather launches with CUDA events and checks the histogram against
the default kernel.

    python tools/sweep_gather.py --config c4 [--orders 3] [--variants tma,win,win-v2,...]

Variants: tma | csr | win (variant 1) | win-lt4 | win-v2 | win-v3, each optionally with
`@a4` = sector-aligned rows (SPB_WIN_ALIGN=4), e.g. `win-v3@a4`.  Unverified variants
are run in this process: wrap the call in `timeout`."""
import argparse
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

ENV = {
    "tma": dict(SPB_GATHER="tma"),
    "csr": dict(SPB_GATHER="csr"),
    "win": dict(SPB_GATHER="win"),
    "win-lt4": dict(SPB_GATHER="win", SPB_WIN_LANE_T="4"),
    "win-v2": dict(SPB_GATHER="win", SPB_WIN_VARIANT="2"),
    "win-v3": dict(SPB_GATHER="win", SPB_WIN_VARIANT="3"),
}
KEYS = ("SPB_GATHER", "SPB_WIN_LANE_T", "SPB_WIN_VARIANT", "SPB_WIN_ALIGN")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--orders", type=int, default=3)
    ap.add_argument("--variants", default="tma,win,win@a4,win-v2,win-v2@a4,win-v3,win-v3@a4")
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
        base, _, opt = name.partition("@")
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(ENV[base])
        if opt == "a4":
            os.environ["SPB_WIN_ALIGN"] = "4"
        rad._tables = None                         # tables depend on the kernel
        t0 = time.time()
        tables = rad._pair_tables(bench.SPEED_OF_SOUND, bench.DT, n_samples)
        torch.cuda.synchronize()
        t_tab = time.time() - t0
        sx = distributed.ShardedExchange(tables, n_samples, dev)
        events = []

        def gather_only(prev, cur, total, b_lo, b_hi, sx=sx, tables=tables):
            t = tables
            c32, sp = _lib.I32(t.dtype), _lib.stream_ptr()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            kind = exchange.gather_kind()
            if t.win_recs is not None and kind != "csr":
                _lib.call("spb_exchange_gather_window", prev, sx.g, t.win_ptr, t.win_recs,
                          sx.cta_order(), t.n_patches, sx.n_alloc, t.n_classes, t.n_dirs,
                          t.n_bands, b_lo, b_hi, sx.j_lo, sx.j_hi, sx.t_pad, sx.ld, sx.pad,
                          exchange.window_arg(t), c32, sp)
            elif kind != "csr":
                _lib.call("spb_exchange_gather_tiled", prev, sx.g, t.ent_ptr, t.recs,
                          sx.cta_order(), t.n_patches, sx.n_alloc, t.n_classes, t.n_dirs,
                          t.n_bands, b_lo, b_hi, sx.j_lo, sx.j_hi, sx.t_pad, sx.ld, sx.pad, c32,
                          sp)
            else:
                _lib.call("spb_exchange_gather", prev, sx.g, t.seg_ptr, t.src, t.wgt, t.dly,
                          t.n_patches, sx.n_alloc, t.n_classes, t.n_dirs, t.n_bands, b_lo, b_hi,
                          sx.j_lo, sx.j_hi, sx.t_pad, sx.ld, sx.pad, c32, sp)
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
