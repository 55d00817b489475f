#!/usr/bin/env python
"""Time the reference's OWN numba kernel `_energy_exchange` (RadiosityFast.py:1073-1145,
serial `njit()`, :1396) next to the oracle's C port on identical inputs, and check that
they agree.  Needs /root/reference, so it runs in the build container only (the GPU box has
no reference tree); its output is committed under profiles/ and relates the `cpu_baseline`
of bench.py (the port, timed on the GPU box) to the numba path that north_star names.

    python tools/time_reference_numba.py > profiles/r02_reference_numba_vs_port.txt
"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
from sparrowpy_b200 import pyfar_shim  # noqa: E402

sys.modules["pyfar"] = pyfar_shim
from ref_import import import_reference  # noqa: E402

sp, RF, geo, ffu, integ = import_reference()
from oracle import oracle as orc  # noqa: E402


def case(name, n, n_dir, n_band, t_len, orders, seed=0):
    rng = np.random.default_rng(seed)
    iu = np.triu_indices(n, 1)
    keep = rng.random(iu[0].size) < 0.7                       # 70 % of the pairs visible
    pairs = np.stack([iu[0][keep], iu[1][keep]], 1).astype(np.int64)
    p = len(pairs)
    c, dt = 343.2, 1e-3
    dist = np.zeros((n, n))
    d = rng.uniform(0.5, 12.0, p)
    dist[pairs[:, 0], pairs[:, 1]] = d
    dist[pairs[:, 1], pairs[:, 0]] = d
    tilde = np.zeros((n, n, n_dir, n_band))
    tilde[pairs[:, 0], pairs[:, 1]] = rng.uniform(0, 2e-3, (p, n_dir, n_band))
    tilde[pairs[:, 1], pairs[:, 0]] = rng.uniform(0, 2e-3, (p, n_dir, n_band))
    p2o = rng.integers(0, n_dir, (n, n))
    e0 = rng.uniform(0, 1, (n, n_dir, n_band))
    d0 = rng.uniform(1.0, 9.0, n)
    args = (t_len, e0, d0, dist, tilde, p2o, c, dt, orders, pairs)
    RF._energy_exchange(t_len, e0, d0, dist, tilde, p2o, c, dt, 1, pairs[:50])   # compile
    t0 = time.perf_counter()
    ref = RF._energy_exchange(*args)
    t_ref = time.perf_counter() - t0
    dsel = np.stack([pairs, pairs[:, ::-1]], 1).reshape(-1, 2)                  # i->j, j->i
    tl = tilde[dsel[:, 0], dsel[:, 1]]
    od = p2o[dsel[:, 0], dsel[:, 1]].astype(np.int64)
    dl = np.repeat((d / c / dt).astype(np.int64), 2)
    t0 = time.perf_counter()
    got = orc.energy_exchange(e0, d0, pairs.astype(np.int32), tl, od, dl, t_len, c, dt, orders,
                              n_threads=1)
    t_port = time.perf_counter() - t0
    x = 2.0 * p * t_len * orders
    err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
    print(f"{name}: N={n} P={p} D={n_dir} B={n_band} T={t_len} K={orders}  "
          f"numba {t_ref:7.2f} s = {x / t_ref:.3e} pair*bin/s   "
          f"C port (1 thread) {t_port:7.2f} s = {x / t_port:.3e}   "
          f"numba/port time = {t_ref / t_port:.2f}   max rel diff {err:.1e}", flush=True)


if __name__ == "__main__":
    import numba
    print(f"# reference numba kernel vs oracle C port, same inputs, one thread each; "
          f"numba {numba.__version__}, numpy {np.__version__}, {os.cpu_count()} host cores "
          f"(build container)")
    orc.build()
    case("diffuse (C4-like: D=1, B=1, T=2000)", 500, 1, 1, 2000, 3)
    case("directional (C2-like: D=16, B=6, T=1000)", 150, 16, 6, 1000, 2)
