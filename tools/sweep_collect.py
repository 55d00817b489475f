"""Time the receiver-collection kernels on a config-3-sized histogram (random data):
`k_collect_partial` (direct) against the ring depths of `k_collect_staged` (the shapes that
were measured -- 16x4, 8x8, 512-thread CTAs -- are recorded under profiles/r02_sweep_collect_*).

    python tools/sweep_collect.py [--bands 16 --patches 40000 --samples 1000 --receivers 64]

One JSON line per variant (ms per launch, FMA TFLOP/s, shared-memory operand TB/s)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bands", type=int, default=16)
    ap.add_argument("--patches", type=int, default=40000)
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--receivers", type=int, default=64)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--variants", default="direct,staged:2,staged:3,staged:4")
    ap.add_argument("--splits", default="0,16,32")
    args = ap.parse_args()
    from sparrowpy_b200 import _lib, exchange
    dev = torch.device("cuda:0")
    tdt = _lib.torch_dtype(_lib.dtype_code(args.dtype))
    b, n, t, r = args.bands, args.patches, args.samples, args.receivers
    pad = 32
    ld = pad + -(-t // 256) * 256
    data = torch.rand((b * n, ld), dtype=tdt, device=dev)
    hist = exchange.EnergyHistogram(data, n, 1, b, t, pad)
    shift = torch.randint(0, t, (r, n), dtype=torch.int32, device=dev)
    scale = torch.rand((r, n, b), dtype=tdt, device=dev)
    rdir = torch.zeros((r, n), dtype=torch.int32, device=dev)
    fma = float(r) * n * b * t
    ref = None
    for var in args.variants.split(","):
        for split in (int(x) for x in args.splits.split(",")):
            if var == "direct" and split:
                continue
            os.environ["SPB_COLLECT"] = var
            try:
                for _ in range(2):
                    out = exchange.collect_mono(hist, rdir, shift, scale, n_split=split or None)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                ev[0].record()
                for _ in range(args.reps):
                    out = exchange.collect_mono(hist, rdir, shift, scale, n_split=split or None)
                ev[1].record()
                torch.cuda.synchronize()
            except Exception as exc:                      # a variant that does not fit
                print(json.dumps({"variant": var, "error": str(exc)[:120]}), flush=True)
                continue
            ms = ev[0].elapsed_time(ev[1]) / args.reps
            if ref is None:
                ref = out
            err = float(((out - ref).abs().max() / ref.abs().max()).item())
            print(json.dumps({"variant": var, "n_split": split or "auto", "ms": round(ms, 3),
                              "fma_tflops": round(2 * fma / ms / 1e9, 2),
                              "operand_tb_s": round(fma * data.element_size() / ms / 1e9, 2),
                              "rel_diff_vs_first": err}), flush=True)


if __name__ == "__main__":
    main()
