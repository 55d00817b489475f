"""Host-side pieces of bench.py that do not need a GPU: the workload description shared by
both arms, the CPU legs (oracle port on bounded samples) and the reference arm's behaviour
on a machine without a CUDA device."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workload_config_is_the_same_object_in_both_arms():
    import bench
    cfg = bench.CONFIGS["c4"]
    a = bench.workload_config(cfg, "c4", 19200, 22291920, 1, 1, "f64")
    b = bench.workload_config(cfg, "c4", 19200, 22291920, 1, 1, "f64")
    assert a == b and a["exchanges_per_etc"] == 2.0 * 22291920 * 2000 * 50
    assert "larger than L2" in a["l2"]
    small = bench.workload_config(bench.CONFIGS["c1"], "c1", 148, 9000, 1, 1, "f64")
    assert "fits L2" in small["l2"]
    assert set(bench.CONFIGS) >= {"c1", "c2", "c3", "c4", "c5"}


def test_cpu_exchange_rate_on_a_synthetic_pair_list():
    """The bounded-sample extrapolation of the CPU baseline: linear in the pair count."""
    import bench
    rng = np.random.default_rng(0)
    n, p = 60, 900
    iu = np.triu_indices(n, 1)
    sel = rng.choice(iu[0].size, p, replace=False)
    pairs = np.stack([iu[0][sel], iu[1][sel]], 1).astype(np.int32)
    inp = dict(n_patches=np.int64(n), pairs=pairs, e0=rng.uniform(0, 1, (n, 1, 1)),
               d0=rng.uniform(1, 5, n), coef=np.ones((1, 1, 1)),
               ff_dir=rng.uniform(0, 1e-3, 2 * p), cls=np.zeros(2 * p, np.int64),
               out_dir=np.zeros(2 * p, np.int64), dist=rng.uniform(0.5, 9.0, p))
    cfg = dict(n_samples=400)
    res = bench.cpu_exchange_rate(inp, cfg, n_threads=1, budget_s=0.5)
    assert res["value"] > 0 and 0 < res["n_sample"] <= p and "extrapolated" in res["sample"]


def test_c3_cpu_leg_runs_on_host_geometry():
    import bench
    import sparrowpy_b200 as sp
    from sparrowpy_b200 import scenes
    rad = sp.DirectionalRadiosityFast.from_polygon(
        [sp.Polygon(*w) for w in scenes.ground_plane(-4, 4, -4, 4)], 0.5)
    srcs, rcvs = bench.grid_points(2, 2, 2.0, 4.0), bench.grid_points(2, 3, 1.5, 4.0)
    assert srcs.shape == (4, 3) and rcvs.shape == (6, 3) and np.abs(srcs[:, :2]).max() < 4
    res = bench.cpu_c3_rate(rad, dict(bench.CONFIGS["c3"], n_samples=200), srcs, rcvs, None,
                            budget_s=1.0)
    assert res["value"] > 0 and res["cores"] == 1 and res["kind"] == "port"


def test_reference_arm_without_a_gpu_reports_unavailable():
    """No CUDA device: the child process that bakes the scene fails loudly, the arm prints
    the contract's `unavailable` line and exits 0 (no CPU fallback for the bake)."""
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    lines = out.stdout.strip().splitlines()
    assert len(lines) == 1, lines          # ONE JSON line on stdout, whatever libraries print
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and "CUDA" in line["unavailable"]
