"""Time the hierarchical visibility kernel on a bench scene: 2-D cells on / off (cells off =
every candidate scan walks the whole y-bin, the behaviour before the cells); the matrices
must be identical.

    python tools/sweep_vis.py --config c4"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import bench
    from sparrowpy_b200 import _lib, bake
    cfg = {k: v for k, v in bench.CONFIGS[args.config].items() if k != "source"}
    rad = bench.build_scene(cfg, "f64", bake=False)
    g = rad._geom()
    n = rad.n_patches
    dev = g["center"].device
    blockers = bake.make_blockers(g["points"], g["normal"])
    tabs = bake.build_groups(blockers.cpu().numpy().reshape(n, -1), rad._patch_to_wall_ids)
    off = tabs[0].copy()
    off.view(np.int32).reshape(len(off), -1)[:, bake._GRP_I["n_bx"]] = 0
    ref = None
    own_in = torch.full((n,), 255, dtype=torch.uint8, device=dev)
    _lib.call("spb_visibility_own_in", g["center"], n, blockers, own_in, _lib.stream_ptr())
    for name, groups, hints in (("cells + own-patch memo + own walls first", tabs[0], True),
                                ("cells", tabs[0], False), ("bins only", off, False)):
        dv = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (groups,) + tabs[1:]]
        vis = torch.empty((n, n), dtype=torch.uint8, device=dev)
        ms = []
        for _ in range(args.reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            _lib.call("spb_visibility_p2p_grouped", g["center"], n, blockers, dv[0], len(groups),
                      dv[1], dv[2], dv[3], dv[4], own_in if hints else None,
                      dv[5] if hints else None, vis, _lib.stream_ptr())
            ev[1].record()
            torch.cuda.synchronize()
            ms.append(ev[0].elapsed_time(ev[1]))
        if ref is None:
            ref = vis.clone()
        print(json.dumps({"config": args.config, "n_patches": n, "tables": name,
                          "ms": [round(x, 2) for x in ms],
                          "visible_pairs": int(vis.sum().item()),
                          "equals_first": bool(torch.equal(vis, ref))}), flush=True)
    time_form_factors(g, ref, args.reps)


def time_form_factors(g, vis, reps=2):
    """Stokes / Nusselt form factors of the visible pairs (k_ff_stokes, k_ff_nusselt)."""
    from sparrowpy_b200 import bake
    pairs = torch.nonzero(vis).to(torch.int32).contiguous()
    ms = []
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        bake.form_factors(g["points"], g["normal"], g["area"], pairs)
        ev[1].record()
        torch.cuda.synchronize()
        ms.append(ev[0].elapsed_time(ev[1]))
    print(json.dumps({"form_factors_ms": [round(x, 2) for x in ms],
                      "pairs": int(pairs.shape[0])}), flush=True)


if __name__ == "__main__":
    main()
