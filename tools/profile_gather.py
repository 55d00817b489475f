#!/usr/bin/env python
"""Lean driver for profiling the stage-1 gather kernels under ncu.

    ncu --set full --import-source on --clock-control none -k regex:k_gather -s 2 -c 1 \\
        -o gpurun_out/gather python tools/profile_gather.py --config c4 --gather win

Bakes the bench scene, builds the pair tables for the chosen kernel and runs a few
reflection orders (one gather + one mix launch each); prints the CUDA-event time of the
gather launches (not a bench number when run under a profiler)."""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--gather", default="win", choices=["tma", "tmem", "win", "csr"])
    ap.add_argument("--orders", type=int, default=4)
    args = ap.parse_args()
    os.environ["SPB_GATHER"] = args.gather
    import torch
    import bench
    from sparrowpy_b200 import _lib, bake, distributed
    cfg = bench.CONFIGS[args.config]
    rad = bench.build_scene(cfg, "f64")
    tables = rad._pair_tables(bench.SPEED_OF_SOUND, bench.DT, cfg["n_samples"], n_shards=1)
    dev = torch.device("cuda", 0)
    sx = distributed.ShardedExchange(tables, cfg["n_samples"], dev)
    e0 = rad._e0_dev.double().contiguous()
    delay0 = bake.delay_bins(rad._d0_dev, bench.SPEED_OF_SOUND, bench.DT)
    events = []
    inner = sx.compute

    def timed(prev, cur, total, b_lo, b_hi):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        inner(prev, cur, total, b_lo, b_hi)
        ev[1].record()
        events.append(ev)

    sx.compute = timed
    sx.init(e0, delay0)
    sx.run(args.orders)
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in events]
    print(f"{args.config} gather={args.gather} records={tables.n_records} "
          f"window={tables.win_w} gather+mix ms per order: "
          + " ".join(f"{x:.2f}" for x in ms), flush=True)


if __name__ == "__main__":
    main()
