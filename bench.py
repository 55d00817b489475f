#!/usr/bin/env python
"""Benchmark of the B200 energy-exchange hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--dtype f64]
    python bench.py --impl reference ...      # CPU arm: oracle port on the host cores

A "step" is one full ETC: `_energy_exchange` with the configuration's reflection
order count on the baked scene (pair tables resident).  Metric: patch-pair x
time-bin exchanges per second, X = 2 * P * T * K_orders per ETC.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

# ---------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------
CONFIGS = {
    # name: scene, patch, (n_az, colatitudes) or None, bands, T, orders, source, receivers
    "c1": dict(scene=("shoebox", (5, 6, 4)), patch=1.0, dirs=None, bands=1, n_samples=1000,
               orders=150, source=(2.0, 2.0, 2.0), receivers=[(2.0, 3.0, 2.0)],
               scattering=1.0, absorption=0.1,
               desc="C1 shoebox 5x6x4 m, 1 m patches, diffuse, 1 band, T=1000, K=150"),
    "c2": dict(scene=("shoebox", (5, 6, 4)), patch=0.2, dirs=(8, (30.0, 60.0)), bands=6,
               n_samples=1000, orders=20, source=(2.0, 2.0, 2.0), receivers=[(2.0, 3.0, 2.0)],
               scattering=0.5, absorption=0.1,
               desc="C2 shoebox 5x6x4 m, 0.2 m patches (N=3700), 16 directions, 6 bands, "
                    "T=1000, K=20"),
    "c3": dict(scene=("plane", 100.0), patch=0.5, dirs=None, bands=1, n_samples=1000, orders=0,
               sources_grid=(4, 4, 2.0), receivers_grid=(8, 8, 1.5), scattering=1.0,
               absorption=0.0, kind="c3",
               desc="C3 infinite-diffuse-plane analogue: 100x100 m ground plane, 0.5 m patches "
                    "(N=40000), 16 sources x 64 receivers, order 0, T=1000"),
    "c4": dict(scene=("canyon", 1.0), patch=1.0, dirs=None, bands=1, n_samples=2000,
               orders=50, source=(60.0, 30.0, 1.5),
               receivers=[(10.0, 30.0, 1.5), (50.0, 28.0, 1.5), (90.0, 32.0, 1.5),
                          (110.0, 30.0, 1.5)],
               scattering=1.0, absorption=0.2,
               desc="C4 street canyon 120x60 m + 5 buildings, 1 m patches (N=19200), "
                    "diffuse, T=2000, K=50"),
    "c5": dict(scene=("city", 1.0), patch=0.5, dirs=(8, (30.0, 60.0)), bands=8,
               n_samples=1000, orders=20, source=(60.0, 37.0, 1.5),
               receivers=[(20.0, 37.0, 1.5), (100.0, 39.0, 1.5)],
               scattering=0.5, absorption=0.2, large=True, first_band_hz=62.5,
               desc="C5 city block 120x75 m + 10 buildings, 0.5 m patches (N=100000), "
                    "16 directions, 8 bands, T=1000, K=20; band-by-band schedule, "
                    "E_total sharded over the GPUs (needs >= 4 GPUs in f64)"),
    "c5s": dict(scene=("city", 0.25), patch=0.5, dirs=(8, (30.0, 60.0)), bands=8,
                n_samples=1000, orders=4, source=(15.0, 9.0, 1.5),
                receivers=[(5.0, 9.0, 1.5), (25.0, 10.0, 1.5)],
                scattering=0.5, absorption=0.2, large=True, first_band_hz=62.5,
                desc="C5-small: city block at 1/4 scale (N~6000), same schedule as C5"),
    # reduced variants for quick checks
    "c2s": dict(scene=("shoebox", (5, 6, 4)), patch=0.5, dirs=(8, (30.0, 60.0)), bands=6,
                n_samples=1000, orders=20, source=(2.0, 2.0, 2.0),
                receivers=[(2.0, 3.0, 2.0)], scattering=0.5, absorption=0.1,
                desc="C2-small shoebox 5x6x4 m, 0.5 m patches (N=592), 16 dirs, 6 bands"),
}
SPEED_OF_SOUND = 343.2
DT = 1e-3


def build_scene(cfg, dtype, bake=True):
    """Bake the scene on the current GPU through the public class API (``bake=False``: set
    the scene up but leave the bake to the caller, e.g. distributed.sharded_bake_tables)."""
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf, scenes
    kind, arg = cfg["scene"]
    walls = (scenes.shoebox(*arg) if kind == "shoebox" else
             scenes.city_block(0, arg) if kind == "city" else
             scenes.ground_plane(-arg / 2, arg / 2, -arg / 2, arg / 2) if kind == "plane" else
             scenes.street_canyon(0, arg))
    rad = sp.DirectionalRadiosityFast.from_polygon([sp.Polygon(*w) for w in walls],
                                                   cfg["patch"], dtype=dtype)
    nb = cfg["bands"]
    freqs = (cfg.get("first_band_hz", 125.0) * 2.0 ** np.arange(nb) if nb > 1
             else np.array([1000.0]))
    if cfg["dirs"] is not None:
        dirs, weights = scenes.hemisphere_directions(*cfg["dirs"])
    else:
        dirs, weights = np.array([[0.0, 0.0, 1.0]]), np.array([1.0])
    brdf = scenes.brdf_from_scattering(dirs, weights, np.full(nb, cfg["scattering"]),
                                       np.full(nb, cfg["absorption"]))
    coords = pf.Coordinates.from_cartesian(dirs, weights=weights)
    rad.set_wall_brdf(np.arange(rad.n_walls), pf.FrequencyData(brdf, freqs), coords, coords)
    air = 1e-4 * 2.0 ** np.arange(nb) if nb > 1 else np.zeros(1)
    rad.set_air_attenuation(pf.FrequencyData(air, freqs))
    if bake:
        rad.bake_geometry()
    if "source" in cfg:
        rad.init_source_energy(pf.Coordinates(*cfg["source"]))
    return rad


def grid_points(nx, ny, z, half):
    """nx x ny points over the central part of a (2 half)^2 plane at height z."""
    xs = (np.arange(nx) + 0.5) / nx * 1.2 * half - 0.6 * half
    ys = (np.arange(ny) + 0.5) / ny * 1.2 * half - 0.6 * half
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    return np.column_stack([gx.ravel(), gy.ravel(), np.full(gx.size, float(z))])


# ---------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                     "--format=csv,noheader,nounits"], capture_output=True, text=True,
                    timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names)
                   if any(len(s) > 2 + k and s[2 + k].lower().startswith("active")
                          for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples else None,
                "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------
# CPU baseline: the oracle port of `_energy_exchange` on a bounded pair sample
# ---------------------------------------------------------------------------
def exchange_inputs(rad):
    """The baked arrays the CPU arms need, as numpy (from a baked + sourced object)."""
    b = rad._baked
    return dict(n_patches=np.int64(rad.n_patches), pairs=b["pairs"].cpu().numpy(),
                e0=rad._energy_init_source, d0=rad._distance_patches_to_source,
                coef=b["coef"].cpu().numpy(), ff_dir=b["ff_dir"].cpu().numpy(),
                cls=b["cls"].cpu().numpy(), out_dir=b["out_dir"].cpu().numpy().astype(np.int64),
                dist=b["dist"].cpu().numpy())


def cpu_exchange_rate(inp, cfg, n_threads, budget_s=12.0, log=None):
    """Time the oracle's `_energy_exchange` (one order) on a random subset of the
    visible pairs and extrapolate to the full pair list (work is exactly linear in
    the number of pairs and orders, RadiosityFast.py:1121-1144).  ``inp``: the arrays of
    :func:`exchange_inputs`.

    Returns dict(value=exchanges/s for the full configuration, ...)."""
    from oracle import oracle as orc
    t_len = cfg["n_samples"]
    pairs = inp["pairs"]
    p_full = pairs.shape[0]
    e0, d0, coef, ff_dir, cls, out_dir = (inp[k] for k in ("e0", "d0", "coef", "ff_dir", "cls",
                                                           "out_dir"))
    delay = np.repeat((inp["dist"] / SPEED_OF_SOUND / DT).astype(np.int64), 2)
    rng = np.random.default_rng(0)

    def run(n_sample, orders=1):
        sel = np.sort(rng.choice(p_full, size=n_sample, replace=False)) if n_sample else \
            np.zeros(0, np.int64)
        dsel = np.stack([2 * sel, 2 * sel + 1], 1).reshape(-1)
        tilde = ff_dir[dsel, None, None] * coef[cls[dsel]]
        t0 = time.perf_counter()
        orc.energy_exchange(e0, d0, pairs[sel], tilde, out_dir[dsel], delay[dsel], t_len,
                            SPEED_OF_SOUND, DT, orders, n_threads=n_threads)
        return time.perf_counter() - t0

    t_fixed = run(0)                       # zeroing + accumulation passes, no pairs
    # Bound the sample a priori: one visible pair costs 2*D*B*T multiply-adds per
    # order and a host core sustains < 4e9 of them per second on this loop, so the
    # probe's timing noise can never blow the budget.
    n_dir, n_band = coef.shape[1], coef.shape[2]
    floor_per_pair = 2.0 * n_dir * n_band * t_len / (4e9 * max(1, n_threads))
    cap = int(max(200, budget_s / floor_per_pair))
    probe = int(min(p_full, max(200, cap // 20)))
    t_probe = run(probe)
    per_pair = max(t_probe - t_fixed, floor_per_pair * probe) / probe
    n_sample = int(min(p_full, cap, max(probe, (budget_s - t_fixed) / per_pair)))
    t_sample = run(n_sample)
    per_pair = max(t_sample - t_fixed, 1e-9) / n_sample
    t_order_full = t_fixed + per_pair * p_full      # one order on the full pair list
    value = 2.0 * p_full * t_len / t_order_full
    if log:
        log(f"cpu baseline: fixed {t_fixed:.2f}s, {n_sample} pairs in {t_sample:.2f}s, "
            f"full order ~{t_order_full:.1f}s")
    return dict(value=value, seconds_per_order=t_order_full, n_sample=n_sample,
                t_sample=t_sample, t_fixed=t_fixed,
                sample=(f"1 reflection order on {n_sample} of {p_full} visible pairs "
                        f"(random, seed 0), full N/D/B/T; extrapolated linearly in pairs "
                        f"and orders; fixed per-order passes {t_fixed:.2f}s included"))


def time_bake(rad, n_pairs):
    """Device times of the two compute-bound bake kernels on the baked scene (north_star
    (1)); their achieved fraction of the FP64 pipe is taken from ncu's instruction counters
    (profiles/), not from assumed op counts."""
    import torch
    from sparrowpy_b200 import bake
    g = rad._geom()

    def timed(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(reps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / reps

    n = rad.n_patches
    blockers = bake.make_blockers(g["points"], g["normal"])
    vis = torch.empty((n, n), dtype=torch.uint8, device=g["center"].device)
    from sparrowpy_b200 import _lib
    vis_ms = timed(lambda: _lib.call("spb_visibility_p2p", g["center"], n, blockers, n, vis,
                                     _lib.stream_ptr()), reps=1)
    # hierarchical variant (the one bake_geometry uses): tables built once, kernel timed
    import numpy as _np
    groups, members, bin_ptr, bin_items, strips, group_of = bake.build_groups(
        blockers.cpu().numpy().reshape(n, -1), rad._patch_to_wall_ids)
    dev = g["center"].device
    gt, mt, bp, bi, sp, og = (torch.from_numpy(_np.ascontiguousarray(a)).to(dev)
                              for a in (groups, members, bin_ptr, bin_items, strips, group_of))
    own_in = torch.empty(n, dtype=torch.uint8, device=dev)
    vis_g = torch.empty_like(vis)

    def grouped():
        _lib.call("spb_visibility_own_in", g["center"], n, blockers, own_in, _lib.stream_ptr())
        _lib.call("spb_visibility_p2p_grouped", g["center"], n, blockers, gt, len(groups), mt,
                  bp, bi, sp, own_in, og, vis_g, _lib.stream_ptr())

    vis_grouped_ms = timed(grouped, reps=1)
    same = bool(torch.equal(vis, vis_g))
    pairs = rad._baked["pairs"]
    ff_ms = timed(lambda: bake.form_factors(g["points"], g["normal"], g["area"], pairs))
    return {
        "visibility_ms": vis_grouped_ms, "visibility_groups": int(len(groups)),
        "visibility_bruteforce_ms": vis_ms, "visibility_grouped_equals_bruteforce": same,
        "pair_wall_tests_per_s": 0.5 * n * (n - 1) * len(groups) / (vis_grouped_ms * 1e-3),
        "form_factor_ms": ff_ms, "form_factor_pairs_per_s": n_pairs / (ff_ms * 1e-3),
        "note": "FP64-pipe fractions of these kernels come from ncu instruction counters "
                "(DFMA/DADD/DMUL executed / duration), see profiles/r02_bake_kernels.txt and "
                "DESIGN.md 3.2-3.3; no nominal op counts are assumed here",
    }


def workload_config(cfg, name, n_patches, n_pairs, n_dir, n_band, dtype):
    """The `config` object of the JSON line: the workload only, identical in both arms
    (kernel / table / parallelism details go under `run`)."""
    esize = 8 if dtype == "f64" else 4
    hist_bytes = n_patches * n_dir * n_band * cfg["n_samples"] * esize
    return {"workload": cfg["desc"], "name": name, "n_patches": int(n_patches),
            "visible_pairs": int(n_pairs), "n_directions": int(n_dir), "n_bands": int(n_band),
            "n_samples": int(cfg["n_samples"]), "reflection_orders": int(cfg["orders"]),
            "exchanges_per_etc": 2.0 * n_pairs * cfg["n_samples"] * cfg["orders"],
            "l2": ("inputs larger than L2 (one energy histogram = "
                   f"{hist_bytes / 1e6:.0f} MB, 3 of them + G are streamed per order)"
                   if hist_bytes > 126e6 * 2 else
                   "working set fits L2 (small config, no flush)")}


def time_pipeline(cfg, dtype):
    """Whole pipeline through the public class, host polygons in, mono ETCs (host) out, with
    wall-clock seconds per stage (CUDA-synchronised): the reference's call sequence
    from_polygon -> set_wall_brdf -> bake_geometry -> init_source_energy ->
    calculate_energy_exchange -> collect_energy_receiver_mono.  Two
    passes on fresh objects: the first one starts from an emptied allocator cache (every buffer
    is a cudaMalloc, whose cost depends on the host), the second is what a process that
    simulates more than one scene pays."""
    first = _pipeline_pass(cfg, dtype)
    again = _pipeline_pass(cfg, dtype)
    again["first_pass"] = {"stages": first["stages"], "total_s": first["total_s"]}
    return again


def _pipeline_pass(cfg, dtype):
    import torch
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import pyfar_shim as pf, scenes
    stages = {}

    def lap(name, t0):
        torch.cuda.synchronize()
        stages[name] = time.perf_counter() - t0
        return time.perf_counter()

    t_all = t0 = time.perf_counter()
    kind, arg = cfg["scene"]
    walls = (scenes.shoebox(*arg) if kind == "shoebox" else scenes.street_canyon(0, arg))
    rad = sp.DirectionalRadiosityFast.from_polygon([sp.Polygon(*w) for w in walls],
                                                   cfg["patch"], dtype=dtype)
    t0 = lap("from_polygon_s", t0)
    nb = cfg["bands"]
    freqs = (cfg.get("first_band_hz", 125.0) * 2.0 ** np.arange(nb) if nb > 1
             else np.array([1000.0]))
    if cfg["dirs"] is not None:
        dirs, weights = scenes.hemisphere_directions(*cfg["dirs"])
    else:
        dirs, weights = np.array([[0.0, 0.0, 1.0]]), np.array([1.0])
    brdf = scenes.brdf_from_scattering(dirs, weights, np.full(nb, cfg["scattering"]),
                                       np.full(nb, cfg["absorption"]))
    coords = pf.Coordinates.from_cartesian(dirs, weights=weights)
    rad.set_wall_brdf(np.arange(rad.n_walls), pf.FrequencyData(brdf, freqs), coords, coords)
    air = 1e-4 * 2.0 ** np.arange(nb) if nb > 1 else np.zeros(1)
    rad.set_air_attenuation(pf.FrequencyData(air, freqs))
    t0 = lap("set_wall_brdf_s", t0)
    rad.bake_geometry()
    t0 = lap("bake_geometry_s", t0)
    rad.init_source_energy(pf.Coordinates(*cfg["source"]))
    t0 = lap("init_source_energy_s", t0)
    rad._pair_tables(SPEED_OF_SOUND, DT, cfg["n_samples"])
    t0 = lap("exchange_tables_s", t0)
    rad.calculate_energy_exchange(SPEED_OF_SOUND, DT, cfg["n_samples"] * DT,
                                  max_reflection_order=cfg["orders"])
    t0 = lap("calculate_energy_exchange_s", t0)
    etc = rad.collect_energy_receiver_mono(
        pf.Coordinates.from_cartesian(np.array(cfg["receivers"], float)))
    t0 = lap("collect_energy_receiver_mono_s", t0)
    return {"stages": stages, "total_s": time.perf_counter() - t_all,
            "mono_etc_checksum": float(np.sum(etc.time)),
            "api": "DirectionalRadiosityFast: host polygons in, mono ETCs at the receivers out "
                   "(a fresh object per pass, nothing cached between passes; first_pass = with "
                   "an empty allocator cache; exchange_tables_s = index tables built on first "
                   "use of calculate_energy_exchange, timed apart)"}


def load_peaks():
    try:
        return json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        return {}


def measure_fma_peak(code, dev, reps=5):
    """FMA-pipe peak of this GPU in the kernel's arithmetic type, measured live with the
    register-only probe kernel of the library (csrc/peak.cu): best of `reps` launches."""
    import ctypes
    import torch
    from sparrowpy_b200 import _lib
    scratch = torch.zeros(16, dtype=torch.float64, device=dev)
    flops = ctypes.c_double(0.0)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    best = float("inf")
    for _ in range(reps + 1):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        _lib.call("spb_fma_peak", _lib.I32(code), 40000 if code == _lib.F64 else 80000,
                  sms * 8, scratch, ctypes.byref(flops), _lib.stream_ptr())
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1))
    return {"tflops": flops.value / (best * 1e-3) / 1e12, "ms": best,
            "how": f"measured live: spb_fma_peak ({'f64' if code == _lib.F64 else 'f32'} FMA "
                   f"chains, {sms * 8} x 256 threads, best of {reps})"}


# ---------------------------------------------------------------------------
_JSON_OUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout, but libraries write there too (NCCL prints its
    version banner with printf when the first communicator is created): keep the real stdout
    for `emit` and send file descriptor 1 -- C stdio included -- to stderr."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    print(json.dumps(line), file=out, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("SPB_BENCH_CONFIG", "c4"),
                    choices=sorted(CONFIGS))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--gather", default=os.environ.get("SPB_GATHER", "tmem"),
                    choices=["tmem", "tma", "csr"],
                    help="stage-1 kernel: tensor-memory windows (f64; f32 falls back to tma), "
                         "TMA-staged tiles or CSR")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bake-to", default=None,
                    help="bake the configuration's scene and write the arrays the CPU arm "
                         "needs to this .npz (used by --impl reference in a child process)")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    quiet_stdout()
    cfg = CONFIGS[args.config]
    os.environ["SPB_GATHER"] = args.gather
    gather = "tma" if (args.gather == "tmem" and args.dtype != "f64") else args.gather
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    def log(msg):
        if args.verbose and rank == 0:
            print(f"[bench] {msg}", file=sys.stderr, flush=True)

    if args.bake_to:
        rad = build_scene(cfg, "f64")
        np.savez(args.bake_to, **exchange_inputs(rad))
        return
    if args.impl == "reference":
        run_reference(args, cfg, rank, world, log)
        return
    import torch

    if cfg.get("large"):
        run_large(args, cfg, rank, world, local_rank, max(args.warmup, 3), log)
        return
    if cfg.get("kind") == "c3":
        run_c3(args, cfg, rank, world, local_rank, max(args.warmup, 3), log)
        return

    import torch.distributed as dist
    from sparrowpy_b200 import _lib, bake, distributed, exchange

    _lib.load()
    _lib.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    t0 = time.time()
    rad = build_scene(cfg, args.dtype)
    torch.cuda.synchronize()
    log(f"baked {args.config}: N={rad.n_patches} P={rad._baked['pairs'].shape[0]} "
        f"in {time.time() - t0:.1f}s")
    n_samples, orders = cfg["n_samples"], cfg["orders"]
    tables = rad._pair_tables(SPEED_OF_SOUND, DT, n_samples, n_shards=world)
    n_pairs = int(rad._baked["pairs"].shape[0])
    bake_info = time_bake(rad, n_pairs) if rank == 0 else None
    n_dir, n_band = tables.n_dirs, tables.n_bands
    x_per_step = 2.0 * n_pairs * n_samples * orders
    code = _lib.dtype_code(args.dtype)
    esize = 8 if code == _lib.F64 else 4

    sx = distributed.ShardedExchange(tables, n_samples, dev)
    e0_dev = rad._e0_dev.to(_lib.torch_dtype(code)).contiguous()
    delay0 = bake.delay_bins(rad._d0_dev, SPEED_OF_SOUND, DT)
    st = torch.cuda.current_stream()
    gather_events = []

    def timed_gather(prev, b_lo, b_hi, j_lo=None, j_hi=None):
        cst = torch.cuda.current_stream()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(cst)
        sx._gather(prev, b_lo, b_hi, j_lo, j_hi)
        ev1.record(cst)
        gather_events.append((ev0, ev1))

    def timed_order_fused(prev, cur, total, b_lo, b_hi):
        cst = torch.cuda.current_stream()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(cst)
        sx._order_fused(prev, cur, total, b_lo, b_hi)
        ev1.record(cst)
        gather_events.append((ev0, ev1))

    sx.gather_fn = timed_gather
    sx.order_fn = timed_order_fused
    fused = sx.fused_order()

    def step():
        sx.init(e0_dev, delay0)
        return sx.run(orders)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    gather_events.clear()
    # gather+mix per (order, band launch) + init (memsets + scatter)
    band_launches = n_band if (world > 1 and sx.comm == "nccl") else 1
    launches_per_step = orders * band_launches * (1 if fused else 2) + 4
    with ClockSampler(local_rank) as clocks:
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if os.environ.get("SPB_BENCH_PROFILE"):     # ncu --profile-from-start off: timed steps only
            torch.cuda.profiler.start()
        ev_a.record(st)
        for _ in range(args.steps):
            step()
        ev_b.record(st)
        barrier()
        if os.environ.get("SPB_BENCH_PROFILE"):
            torch.cuda.profiler.stop()
        elapsed_ms = ev_a.elapsed_time(ev_b)
    gather_ms = [a.elapsed_time(b) for a, b in gather_events]
    if world > 1:
        tmax = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    ms_per_step = elapsed_ms / args.steps
    value = x_per_step / (ms_per_step * 1e-3)

    # -- roofline of the dominant kernel (stage-1 gather), this rank's share ----------
    peaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    share = 1.0 / world     # shards are balanced by dealing the receiver tiles round-robin
    # SURVEY 8(d) algorithmic bytes per pair.bin exchange = B*(1+2D)*sizeof (the reference's
    # dense-tilde formulation); one gather launch processes this rank's directed pairs x T
    alg_bytes_launch = (2.0 * n_pairs * share * n_samples * n_band * (1 + 2 * n_dir) * esize
                        / band_launches)
    gather_avg_ms = float(np.mean(gather_ms)) if gather_ms else float("nan")
    alg_gbs = alg_bytes_launch / (gather_avg_ms * 1e-3) / 1e9
    # FMAs of the factored kernel per launch.  useful: one per (kept directed pair, band,
    # bin); executed: what the records make the pipe do (8 receiver slots per record, empty
    # slots multiply by zero, the time axis is rounded up to the CTA's bin count)
    t_exec = -(-sx.t_pad // 512) * 512 if gather == "tmem" else sx.t_pad
    useful_fma = float(tables.src.numel()) * share * n_band * n_samples / band_launches
    exec_fma = (float(tables.n_records) * (1.0 if world == 1 else share) * 8 * n_band * t_exec
                / band_launches) if gather != "csr" else useful_fma * sx.t_pad / n_samples
    if world > 1 and tables.n_records:          # this rank's records, counted exactly
        ptr = tables.tile_ptr.view(-1)
        n_blocks = -(-tables.n_patches // 8)
        cnt = (ptr[1:] - ptr[:-1]).view(tables.n_classes, n_blocks)
        mine = int(cnt[:, sx.j_lo // 8:-(-sx.j_hi // 8)].sum().item())
        exec_fma = float(mine) * 8 * n_band * t_exec / band_launches
    exec_tflops = 2.0 * exec_fma / (gather_avg_ms * 1e-3) / 1e12
    pipe = measure_fma_peak(code, dev)
    traffic = None
    try:
        tr = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))
        traffic = tr.get(f"{args.config}/{args.dtype}/{gather}", {}).get("bytes")
        if traffic is not None and world > 1:
            traffic = None                      # captured at 1 GPU only
    except Exception:  # noqa: BLE001
        pass
    kernel = {"tmem": "k_gather_tmem", "tma": "k_gather_tma", "csr": "k_gather"}[gather]
    if fused:
        kernel += " (stage 2 and the delivery of E_k fused into its epilogue)"
    roofline = {
        # the gather is bound by its FMA pipe and the operand path that feeds it, not by HBM
        # (DRAM traffic is a few % of the HBM peak, see `traffic` and profiles/); the
        # north-star's HBM formulation is kept below as `hbm_algorithmic`
        "bound": "fp64" if code == _lib.F64 else "fp32",
        "achieved": exec_tflops, "peak": pipe["tflops"], "unit": "TFLOP/s",
        "frac": exec_tflops / pipe["tflops"], "traffic": traffic,
        "peak_source": pipe["how"], "kernel": kernel,
        "avg_launch_ms": gather_avg_ms, "launches_timed": len(gather_ms),
        "share_of_step": sum(gather_ms) / max(elapsed_ms, 1e-9),
        "executed_fma_per_launch": exec_fma, "useful_fma_per_launch": useful_fma,
        "useful_tflops": 2.0 * useful_fma / (gather_avg_ms * 1e-3) / 1e12,
        "frac_useful": 2.0 * useful_fma / (gather_avg_ms * 1e-3) / 1e12 / pipe["tflops"],
        "record_slot_fill": float(tables.src.numel()) / max(1.0, 8.0 * tables.n_records),
        "fma_pipe_nominal_tflops": 37.0 if code == _lib.F64 else 75.0,
        "hbm_algorithmic": {
            "bytes_per_launch": alg_bytes_launch, "achieved": alg_gbs, "peak": hbm_peak,
            "unit": "GB/s", "frac": alg_gbs / hbm_peak, "frac_nominal_8tbs": alg_gbs / 8000.0,
            "peak_source": ("measured (MEASURED_PEAKS.json hbm_gbs)" if peaks
                            else "fallback 6650 GB/s"),
            "note": "SURVEY 8(d): bytes of the reference's dense-tilde formulation, "
                    "B*(1+2D)*sizeof per pair.bin; the factored kernel moves far fewer, so "
                    "this fraction can exceed 1 and is not a utilisation"},
        "dram_frac_of_hbm_peak": (traffic / (gather_avg_ms * 1e-3) / 1e9 / hbm_peak
                                  if traffic else None)}

    # -- end to end with HOST buffers: E0 / distances in pinned host memory -> device, K
    # orders, full ETC -> pinned host.  One GPU: through the public operator
    # exchange.energy_exchange_host; sharded: every rank uploads the inputs (each needs all
    # senders), runs its receiver shard, rank 0 reads the gathered ETC back
    tdt = _lib.torch_dtype(code)
    e0_host = rad._e0_dev.cpu().pin_memory()
    d0_host = rad._d0_dev.cpu().pin_memory()
    n, d, b = e0_host.shape
    out_host = (torch.empty((n, d, b, n_samples), dtype=tdt).pin_memory()
                if rank == 0 else None)
    if world == 1:
        ws = exchange.ExchangeWorkspace(tables, n_samples, dev)

        def e2e_step():
            exchange.energy_exchange_host(tables, e0_host, d0_host, SPEED_OF_SOUND, DT,
                                          n_samples, orders, out_host, workspace=ws)
        api = ("sparrowpy_b200.exchange.energy_exchange_host (host E0/d0 in pinned memory "
               "-> device, K orders, full ETC -> pinned host)")
    else:
        sx.gather_fn, sx.order_fn = sx._gather, sx._order_fused

        def e2e_step():
            e0 = e0_host.to(dev, non_blocking=True).to(tdt)
            d0 = d0_host.to(dev, non_blocking=True)
            sx.init(e0, bake.delay_bins(d0, SPEED_OF_SOUND, DT))
            hist = sx.run(orders)
            if rank == 0:
                out_host.copy_(hist.dense())
        api = ("host E0/d0 (pinned) -> every rank's device, distributed.ShardedExchange "
               "init + run (E_total all-gathered), full ETC -> rank 0's pinned host")
    for _ in range(2):
        e2e_step()
    barrier()
    n_e2e = max(2, min(args.steps, 5))
    t1 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t1) / n_e2e
    if world > 1:
        tmax = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e = {"value": x_per_step / e2e_s, "unit": "pair*bin exchanges/s",
           "h2d_bytes_per_step": int(e0_host.numel() * 8 + d0_host.numel() * 8) * world,
           "d2h_bytes_per_step": int(n * d * b * n_samples * esize),
           "ms_per_step": e2e_s * 1e3, "api": api}

    # -- result evidence: checksum of the ETC and, when sharded, bit-equality with the
    # single-GPU schedule run on rank 0's device with the same tables
    result = None
    if rank == 0:
        result = {"etc_checksum": float(out_host.double().sum()),
                  "etc_abs_max": float(out_host.abs().max())}
        if world > 1:
            one = distributed.ShardedExchange(tables, n_samples, dev, local=True)
            one.init(e0_host.to(dev).to(tdt), bake.delay_bins(d0_host.to(dev), SPEED_OF_SOUND, DT))
            single = one.run(orders).dense()
            result["equals_single_gpu_bitwise"] = bool(
                torch.equal(single.cpu(), out_host))
            result["single_gpu_checksum"] = float(single.double().sum())
            del one, single
    barrier()

    # -- CPU baseline on the host cores (rank 0, N = 1 only) ----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = cpu_exchange_rate(exchange_inputs(rad), cfg, n_threads=1, log=log)
        cpu = {"value": res["value"], "unit": "pair*bin exchanges/s", "cores": 1,
               "kind": "port", "sample": res["sample"],
               "seconds_per_etc": res["seconds_per_order"] * orders}

    pipeline = None
    comm = sx.comm
    if rank == 0 and world == 1:
        del sx
        torch.cuda.empty_cache()
        pipeline = time_pipeline(cfg, args.dtype)
    if rank == 0:
        line = {
            "metric": "patch-pair*time-bin exchanges/s (energy exchange, s per ETC alongside)",
            "value": value, "unit": "pair*bin exchanges/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "seconds_per_etc": ms_per_step * 1e-3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic", "gpu_launches": int(launches_per_step * args.steps),
            "config": workload_config(cfg, args.config, rad.n_patches, n_pairs, n_dir, n_band,
                                      args.dtype),
            "run": {"gather": gather, "kernel": kernel,
                    "directed_pairs_kept": int(tables.src.numel()),
                    "tile_records": int(tables.n_records),
                    "record_window": int(tables.win_w),
                    "parallelism": (f"receiver shards x{world}, exchange: {comm}"
                                    if world > 1 else "1 GPU")},
            "clocks": clocks.summary(), "roofline": roofline,
        }
        if bake_info:
            line["bake"] = bake_info
        line["e2e"] = e2e
        line["result"] = result
        if pipeline:
            line["pipeline"] = pipeline
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_c3(args, cfg, rank, world, local_rank, warmup, log):
    """Config 3: one ground plane (coplanar patches never see each other, P = 0), many
    sources and receivers, reflection order 0 -- the reference's infinite-diffuse-plane
    test (tests/test_DRadiosityFast_infinite_diffuse_plane.py:59-90), where it loops over
    receivers in Python (RadiosityFast.py:711) and allows one source (:450-451).  A step =
    source energies of all S sources (one launch), order-0 histograms, and the mono ETC of
    all R receivers.  Metric: (source, receiver, patch, band, bin) contributions per second,
    S*R*N*B*T per step -- the FMAs of `_collect_receiver_energy` (RadiosityFast.py:1148-1185).
    Receivers are sharded over the ranks (no collective in the data path)."""
    import torch
    import torch.distributed as dist
    from sparrowpy_b200 import _lib, bake, exchange, pyfar_shim as pf

    _lib.load()
    _lib.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.time()
    rad = build_scene(cfg, args.dtype)
    torch.cuda.synchronize()
    n = rad.n_patches
    half = cfg["scene"][1] / 2
    srcs = grid_points(*cfg["sources_grid"], half)
    rcvs_all = grid_points(*cfg["receivers_grid"], half)
    n_src, n_rcv_all = len(srcs), len(rcvs_all)
    rcvs = rcvs_all[rank::world]
    n_samples = cfg["n_samples"]
    log(f"baked c3: N={n} P={rad._baked['pairs'].shape[0]} in {time.time() - t0:.1f}s; "
        f"{n_src} sources, {len(rcvs)} of {n_rcv_all} receivers on this rank")
    code = _lib.dtype_code(args.dtype)
    esize = 8 if code == _lib.F64 else 4
    tdt = _lib.torch_dtype(code)
    g = rad._geom()
    rad._source_energy(srcs[:1])              # installs the default BRDF / air tables
    vi, vo, brdf, bidx = rad._brdf_tables()
    air = torch.from_numpy(np.real(rad._air_attenuation).astype(float)).to(dev)
    vi_d, vo_d = torch.from_numpy(vi).to(dev), torch.from_numpy(vo).to(dev)
    brdf_d, bidx_d = torch.from_numpy(brdf).to(dev), torch.from_numpy(bidx).to(dev)
    src_d = torch.from_numpy(srcs).to(dev)
    rcv_d = torch.from_numpy(np.ascontiguousarray(rcvs)).to(dev)
    n_band = int(air.shape[0])
    empty = torch.zeros(0, dtype=torch.int64, device=dev)
    tables = exchange.build_pair_tables(
        empty, empty, torch.zeros(0, dtype=torch.float64, device=dev), empty, empty, empty,
        torch.ones((1, vo.shape[1], n_band), dtype=torch.float64, device=dev), n, n_samples,
        args.dtype)
    ev = {k: [] for k in ("source", "init", "factors", "collect")}
    last = {}

    def timed(key, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        ev[key].append((e0, e1))
        return out

    def step():
        svis = bake.visibility_pt2p(src_d, g["center"], g["walls_normal"], g["walls_points"])
        d0, e0 = timed("source", lambda: bake.source_energy_batch(
            src_d, g["center"], g["points"], svis, air, g["wall_ids"], vi_d, brdf_d, bidx_d,
            vo.shape[1]))
        hist = timed("init", lambda: exchange.energy_exchange(
            tables, e0, bake.delay_bins(d0.reshape(-1), SPEED_OF_SOUND, DT).view(n_src, n),
            n_samples, 0))
        rvis = bake.visibility_pt2p(rcv_d, g["center"], g["walls_normal"], g["walls_points"])
        rt = timed("factors", lambda: bake.receiver_factors(
            rcv_d, g["center"], g["points"], rvis, air, g["wall_ids"], vo_d, SPEED_OF_SOUND, DT,
            n_samples))
        last["hist"] = hist
        return timed("collect", lambda: exchange.collect_mono(
            hist, rt["rdir"], rt["shift"], rt["scale"]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    st = torch.cuda.current_stream()
    for _ in range(warmup):
        step()
    barrier()
    for v in ev.values():
        v.clear()
    with ClockSampler(local_rank) as clocks:
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if os.environ.get("SPB_BENCH_PROFILE"):     # ncu --profile-from-start off: timed steps only
            torch.cuda.profiler.start()
        ev_a.record(st)
        for _ in range(args.steps):
            mono = step()
        ev_b.record(st)
        barrier()
        if os.environ.get("SPB_BENCH_PROFILE"):
            torch.cuda.profiler.stop()
        elapsed_ms = ev_a.elapsed_time(ev_b)
    stage_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in ev.items()}
    if world > 1:
        tmax = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    ms_per_step = elapsed_ms / args.steps
    x_per_step = float(n_src) * n_rcv_all * n * n_band * n_samples
    value = x_per_step / (ms_per_step * 1e-3)

    # end to end through the class: host coordinates in, TimeData (host) out
    src_c = pf.Coordinates.from_cartesian(srcs)
    rcv_c = pf.Coordinates.from_cartesian(rcvs)

    def e2e_step():
        rad.init_source_energy_batch(src_c)
        rad.calculate_energy_exchange(SPEED_OF_SOUND, DT, n_samples * DT, max_reflection_order=0,
                                      recalculate=True)
        return rad.collect_energy_receiver_mono(rcv_c).time

    e2e_step()
    barrier()
    n_e2e = max(2, min(args.steps, 5))
    t1 = time.perf_counter()
    for _ in range(n_e2e):
        etc = e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t1) / n_e2e
    if world > 1:
        tmax = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    same = float(np.max(np.abs(etc - mono.double().cpu().numpy())) / np.max(np.abs(etc)))

    # roofline of the collection kernel: one 8-byte histogram read per contribution
    peaks = load_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    fma = float(n_src) * len(rcvs) * n * n_band * n_samples
    alg_bytes = fma * esize
    compulsory = float(n_src) * n_band * n * vo.shape[1] * n_samples * esize
    pipe = measure_fma_peak(code, dev)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_c3_rate(rad, cfg, srcs, rcvs_all, log)
    if rank == 0:
        c_ms = stage_ms["collect"]
        clk = clocks.summary()
        staged = exchange.collect_kind(last["hist"], len(rcvs))[0] == "staged"
        # k_collect_staged reads one shared-memory operand per FMA: its roof is the SM's
        # 128 B/clk shared-memory path (a quarter of the FP64 pipe), at the SM clock sampled
        # during the timed region.  k_collect_partial reads the operand from L2/HBM.
        lsu_peak = 128.0 * 148 * float(clk.get("sm_mhz") or 1965.0) * 1e6 / 1e9
        achieved = alg_bytes / (c_ms * 1e-3) / 1e9
        peak = lsu_peak if staged else hbm_peak
        roof = {
            "bound": "lsu" if staged else "hbm",
            "kernel": "k_collect_staged" if staged else "k_collect_partial",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": c3_traffic("k_collect_staged" if staged else "k_collect_partial",
                                  args.dtype, world),
            "peak_source": ("128 B/clk/SM shared-memory operand path x 148 SMs x the SM clock "
                            "sampled in the timed region" if staged
                            else "measured (MEASURED_PEAKS.json hbm_gbs)"),
            "avg_launch_ms": c_ms, "share_of_step": c_ms / ms_per_step,
            "algorithmic_bytes_per_launch": alg_bytes,
            "compulsory_hbm_bytes_per_launch": compulsory,
            "compulsory_hbm_frac_of_peak": compulsory / (c_ms * 1e-3) / 1e9 / hbm_peak,
            "fma_tflops": 2.0 * fma / (c_ms * 1e-3) / 1e12,
            "fma_frac_of_measured_pipe": 2.0 * fma / (c_ms * 1e-3) / 1e12 / pipe["tflops"],
            "fma_pipe_measured_tflops": pipe["tflops"],
            "note": "algorithmic bytes = one histogram element per (source, receiver, patch, "
                    "band, bin) contribution (SURVEY 8d); the staged kernel reads each of them "
                    "from shared memory (the row of a (source, patch) is fetched once per "
                    "group of 8 receivers), so the operand path, not HBM, is the roof"}
        line = {
            "metric": "source*receiver*patch*band*bin contributions/s (order-0 ETCs at all "
                      "receivers; BASELINE config 3)",
            "value": value, "unit": "contributions/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "gpu_launches": int(args.steps * (8 + 2 * n_src)),
            "config": {"workload": cfg["desc"], "name": args.config, "n_patches": n,
                       "visible_pairs": 0, "n_sources": n_src, "n_receivers": n_rcv_all,
                       "n_directions": int(vo.shape[1]), "n_bands": n_band,
                       "n_samples": n_samples, "reflection_orders": 0,
                       "contributions_per_step": x_per_step,
                       "l2": "inputs larger than L2 (order-0 histograms of all sources = "
                             f"{compulsory / 1e9:.1f} GB)"},
            "run": {"parallelism": f"receivers sharded x{world}" if world > 1 else "1 GPU",
                    "stage_ms": stage_ms},
            "clocks": clk,
            "roofline": roof,
            "e2e": {"value": x_per_step / e2e_s, "unit": "contributions/s",
                    "h2d_bytes_per_step": int((srcs.size + rcvs.size) * 8),
                    "d2h_bytes_per_step": int(etc.size * 8), "ms_per_step": e2e_s * 1e3,
                    "api": "init_source_energy_batch + calculate_energy_exchange(order 0) + "
                           "collect_energy_receiver_mono (host coordinates in, host ETCs out)"},
            "result": {"etc_checksum": float(etc.sum()), "class_vs_operator_rel_diff": same},
        }
        if cpu:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def c3_traffic(kernel, dtype, world):
    """DRAM bytes per launch of the collection kernel from the ncu capture under profiles/
    (1 GPU, all receivers), None when there is none."""
    if world > 1:
        return None
    try:
        tr = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))
        return tr.get(f"c3/{dtype}/{kernel}", {}).get("bytes")
    except OSError:
        return None


def cpu_c3_rate(rad, cfg, srcs, rcvs, log, budget_s=15.0):
    """The oracle's order-0 path (source energy, initial histogram, receiver collection:
    universal.py:98-160, RadiosityFast.py:1037-1070, :1148-1185) for one source and a bounded
    number of receivers, serial like the reference; extrapolated linearly to S x R."""
    from oracle import oracle as orc
    cen, pts, ids = rad.patches_center, rad.patches_points, rad._patch_to_wall_ids
    wp, wn = rad._walls_points, rad._walls_normal
    one = np.array([[[0.0, 0.0, 1.0]]])
    brdf = np.full((1, 1, 1, 1), np.pi)
    air = np.zeros(1)
    t0 = time.perf_counter()
    svis = orc.visibility_pt2p(srcs[0], cen, wn, wp)
    e0b, d0 = orc.source_energy(srcs[0], cen, pts, svis, air)
    e0 = orc.add_directional(e0b, srcs[0], cen, ids, one, one, brdf, np.zeros(1, np.int64))
    etc = orc.init_energy(e0, d0, cfg["n_samples"], SPEED_OF_SOUND, DT)
    t_src = time.perf_counter() - t0
    n_r, t_rcv = 0, 0.0
    while n_r < len(rcvs) and t_src + t_rcv < budget_s:
        t0 = time.perf_counter()
        r = rcvs[n_r]
        v = orc.visibility_pt2p(r, cen, wn, wp)
        f = orc.receiver_factor(r, pts, v)
        k = orc.receiver_dir_index(cen, r, one, ids)
        orc.collect_receiver(etc, r, cen, f, k, air, SPEED_OF_SOUND, DT)
        t_rcv += time.perf_counter() - t0
        n_r += 1
    t_full = len(srcs) * (t_src + t_rcv / n_r * len(rcvs))
    x = float(len(srcs)) * len(rcvs) * rad.n_patches * cfg["n_samples"]
    if log:
        log(f"cpu c3: source stage {t_src:.2f}s, {n_r} receivers {t_rcv:.2f}s")
    return {"value": x / t_full, "unit": "contributions/s", "cores": 1, "kind": "port",
            "seconds_per_step": t_full,
            "sample": f"oracle order-0 path for 1 of {len(srcs)} sources and {n_r} of "
                      f"{len(rcvs)} receivers (coplanar scene: no patch pairs), extrapolated "
                      "linearly to all sources and receivers"}


def run_large(args, cfg, rank, world, local_rank, warmup, log):
    """Large-scene arm (BASELINE config 5): the histograms of the whole scene do not fit
    one GPU, so the exchange runs band by band (distributed.BandwiseExchange), every rank
    keeps `E_total` of its own receivers only, the pair tables are built per receiver
    shard, and the result a user reads is the mono ETC at the receivers (all-reduced) --
    the full (N, D, B, T) histogram never exists.  Same metric and timing rules as the
    main arm; no CPU baseline (run at N > 1) and no brute-force visibility timing."""
    import torch
    import torch.distributed as dist
    from sparrowpy_b200 import _lib, bake, distributed

    _lib.load()
    _lib.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    t0 = time.time()
    rad = build_scene(cfg, args.dtype, bake=False)
    n_samples, orders = cfg["n_samples"], cfg["orders"]
    # the bake itself is sharded: every rank evaluates its rows of the visibility matrix and
    # the pairs found there, then the directed pairs travel to the owner of their receiver
    # (distributed.sharded_bake_tables); no rank holds the whole matrix or pair list
    torch.cuda.reset_peak_memory_stats(dev)
    tables, n_pairs = distributed.sharded_bake_tables(rad, SPEED_OF_SOUND, DT, n_samples)
    torch.cuda.synchronize()
    bake_s = time.time() - t0
    bake_peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
    log(f"sharded bake {args.config}: N={rad.n_patches} P={n_pairs} in {bake_s:.1f}s, "
        f"peak {bake_peak_gb:.1f} GB on rank 0")
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats(dev)
    code = _lib.dtype_code(args.dtype)
    esize = 8 if code == _lib.F64 else 4
    n_dir, n_band = tables.n_dirs, tables.n_bands
    x_per_step = 2.0 * n_pairs * n_samples * orders
    bx = distributed.BandwiseExchange(tables, n_samples, dev,
                                      band_block=int(cfg.get("band_block", 1)))
    sx = bx.sx
    e0_dev = rad._e0_dev.to(_lib.torch_dtype(code)).contiguous()
    delay0 = bake.delay_bins(rad._d0_dev, SPEED_OF_SOUND, DT)
    gather_events = []
    inner = sx.compute

    def timed_order(prev, cur, total, b_lo, b_hi):
        # the gather is the first launch of the local step; bracket the whole step and
        # the mix separately would need a second event pair -- the step is >95 % gather
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cst = torch.cuda.current_stream()
        ev0.record(cst)
        inner(prev, cur, total, b_lo, b_hi)
        ev1.record(cst)
        gather_events.append((ev0, ev1))

    sx.compute = timed_order
    g = rad._geom()
    rcv = torch.tensor(cfg["receivers"], dtype=torch.float64, device=dev)
    rvis = bake.visibility_pt2p(rcv, g["center"], g["walls_normal"], g["walls_points"])
    _, vo, _, _ = rad._brdf_tables()
    air = torch.from_numpy(np.real(rad._air_attenuation).astype(float)).to(dev)
    rt = bake.receiver_factors(rcv, g["center"], g["points"], rvis, air, g["wall_ids"],
                               torch.from_numpy(vo).to(dev), SPEED_OF_SOUND, DT, n_samples)

    def step():
        return bx.run(e0_dev, delay0, orders)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    st = torch.cuda.current_stream()
    for _ in range(warmup):
        step()
    barrier()
    gather_events.clear()
    with ClockSampler(local_rank) as clocks:
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev_a.record(st)
        for _ in range(args.steps):
            hist = step()
        ev_b.record(st)
        barrier()
        elapsed_ms = ev_a.elapsed_time(ev_b)
    step_ms = [a.elapsed_time(b) for a, b in gather_events]
    if world > 1:
        tmax = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    ms_per_step = elapsed_ms / args.steps
    value = x_per_step / (ms_per_step * 1e-3)

    # end to end: host E0 / distances in, mono ETCs at the receivers out
    e0_host = rad._e0_dev.cpu().pin_memory()
    d0_host = rad._d0_dev.cpu().pin_memory()
    mono_host = torch.empty((len(cfg["receivers"]), n_band, n_samples),
                            dtype=_lib.torch_dtype(code)).pin_memory()

    def e2e_step():
        e0 = e0_host.to(dev, non_blocking=True).to(_lib.torch_dtype(code))
        d0 = d0_host.to(dev, non_blocking=True)
        h = bx.run(e0, bake.delay_bins(d0, SPEED_OF_SOUND, DT), orders)
        mono_host.copy_(h.collect_mono(rt["rdir"], rt["shift"], rt["scale"]))

    e2e_step()
    barrier()
    t1 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 3))
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t1) / n_e2e
    if world > 1:
        tmax = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    launches = len(step_ms)                     # local steps (one band block, one order)
    alg_bytes = (2.0 * n_pairs / world * n_samples * bx.band_block * (1 + 2 * n_dir) * esize)
    avg_ms = float(np.mean(step_ms)) if step_ms else float("nan")
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    if rank == 0:
        line = {
            "metric": "patch-pair*time-bin exchanges/s (energy exchange, s per ETC alongside)",
            "value": value, "unit": "pair*bin exchanges/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "seconds_per_etc": ms_per_step * 1e-3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic", "gpu_launches": int(2 * launches + 4 * n_band * args.steps),
            "config": {"workload": cfg["desc"], "name": args.config,
                       "n_patches": rad.n_patches, "visible_pairs": n_pairs,
                       "directed_pairs_kept_this_rank": int(tables.src.numel()),
                       "tile_records_this_rank": int(tables.n_records),
                       "gather": args.gather, "n_directions": n_dir, "n_bands": n_band,
                       "n_samples": n_samples, "reflection_orders": orders,
                       "band_block": bx.band_block, "exchanges_per_etc": x_per_step,
                       "l2": "inputs larger than L2",
                       "parallelism": (f"receiver shards x{world}, exchange: {sx.comm}, "
                                       "bands sequential, E_total sharded")},
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                         "kernel": "k_gather_tmem + k_mix (one band block, one order)",
                         "avg_launch_ms": avg_ms, "launches_timed": launches,
                         "share_of_step": sum(step_ms) / max(elapsed_ms, 1e-9),
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "see the main arm: algorithmic bytes of the reference's "
                                 "dense formulation, frac can exceed 1"},
            "e2e": {"value": x_per_step / e2e_s, "unit": "pair*bin exchanges/s",
                    "h2d_bytes_per_step": int(e0_host.numel() * 8 + d0_host.numel() * 8),
                    "d2h_bytes_per_step": int(mono_host.numel() * esize),
                    "ms_per_step": e2e_s * 1e3,
                    "api": "host E0/d0 -> BandwiseExchange.run -> ShardedHistogram."
                           "collect_mono (all-reduce) -> pinned host"},
            "result_checksum": float(mono_host.double().sum()),
            "bake": {"sharded_bake_s": bake_s, "peak_memory_gb_rank0": bake_peak_gb,
                     "note": "visibility rows / form factors / direction indices split over "
                             "the ranks, directed pairs routed to the receiver's owner by one "
                             "all-to-all (distributed.sharded_bake_tables)"},
            "exchange_peak_memory_gb_rank0": torch.cuda.max_memory_allocated(dev) / 1e9,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args, cfg, rank, world, log):
    """CPU arm: the oracle port of the reference's `_energy_exchange` on the host
    cores (all threads), bounded sample per step.  The scene's baked arrays are INPUTS of
    this arm: they are produced by a child process (`bench.py --bake-to file`, the CUDA
    bake -- the CPU oracle would need hours for the visibility of 19 200 patches) and read
    back as numpy, so that this process never loads the CUDA library and nothing of the
    GPU path is inside or beside the timed region."""
    if rank != 0:
        return
    import tempfile
    from oracle import oracle as orc
    orc.build()
    if cfg.get("kind") == "c3":                 # no patch pairs: host geometry is all it needs
        import sparrowpy_b200 as sp
        from sparrowpy_b200 import scenes
        half = cfg["scene"][1] / 2
        rad = sp.DirectionalRadiosityFast.from_polygon(
            [sp.Polygon(*w) for w in scenes.ground_plane(-half, half, -half, half)], cfg["patch"])
        srcs = grid_points(*cfg["sources_grid"], half)
        rcvs = grid_points(*cfg["receivers_grid"], half)
        res = cpu_c3_rate(rad, cfg, srcs, rcvs, log, budget_s=20.0)
        emit({
            "impl": "reference", "metric": "source*receiver*patch*band*bin contributions/s "
            "(order-0 ETCs at all receivers; BASELINE config 3)", "value": res["value"],
            "unit": "contributions/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": cfg["desc"], "name": args.config},
            "cpu_baseline": res,
            "e2e": {"value": res["value"], "unit": "contributions/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}})
        return
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "baked.npz")
        env = {k: v for k, v in os.environ.items()
               if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE")}
        child = subprocess.run([sys.executable, os.path.abspath(__file__), "--config",
                                args.config, "--bake-to", path], env=env,
                               capture_output=True, text=True)
        if child.returncode != 0 or not os.path.exists(path):
            why = (child.stderr.strip().splitlines() or ["bake failed"])[-1][:200]
            emit({"impl": "reference",
                  "unavailable": f"scene baking needs the CUDA path: {why}"})
            return
        inp = dict(np.load(path))
    n_patches = int(inp["n_patches"])
    n_pairs = int(inp["pairs"].shape[0])
    n_dir, n_band = int(inp["coef"].shape[1]), int(inp["coef"].shape[2])
    # thread count set explicitly (torchrun exports OMP_NUM_THREADS=1; the oracle's
    # `num_threads` clause does not depend on it): every core this process may run on
    threads = max(1, len(os.sched_getaffinity(0)))
    orders = cfg["orders"]
    x_per_step = 2.0 * n_pairs * cfg["n_samples"] * orders
    budget = max(2.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    res = None
    for k in range(args.warmup + args.steps):
        res = cpu_exchange_rate(inp, cfg, n_threads=threads, budget_s=budget, log=log)
        if k >= args.warmup:
            vals.append(res["value"])
    value = float(np.mean(vals)) if vals else float("nan")
    one = cpu_exchange_rate(inp, cfg, n_threads=1, budget_s=6.0, log=log)
    line = {
        "impl": "reference",
        "metric": "patch-pair*time-bin exchanges/s (energy exchange, s per ETC alongside)",
        "value": value, "unit": "pair*bin exchanges/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": x_per_step / value * 1e3 if value == value else None,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(cfg, args.config, n_patches, n_pairs, n_dir, n_band,
                                  "f64"),
        "cpu_baseline": {"value": value, "unit": "pair*bin exchanges/s", "cores": threads,
                         "kind": "port", "sample": res["sample"] if res else "",
                         "value_1core": one["value"],
                         "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "e2e": {"value": value, "unit": "pair*bin exchanges/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "oracle port of _energy_exchange (time-sliced over all host threads, "
                "bit-identical to the serial reference order); the reference itself is "
                "single-threaded here (RadiosityFast.py:1396) -- value_1core is the like-for-"
                "like figure; the scene's baked arrays are inputs, produced by a child process "
                "(CUDA bake) -- this process never loads the CUDA library",
    }
    emit(line)


if __name__ == "__main__":
    main()
