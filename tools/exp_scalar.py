"""Development aid: where a scalar likelihood call spends its time (Python layer vs the C-ABI call)."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, quick_bench as qb
from mcalf_b200 import capi
for cfg in (4, 2, 1):
    g = qb.make(cfg)
    p = g._scale_cube_pc(np.random.default_rng(0).random(g.ndim))
    for _ in range(50): g.lnlhood_worker(p)
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n): g.lnlhood_worker(p)
    t_py = (time.perf_counter() - t0) / n
    row = np.ascontiguousarray(p[None, :]); out = np.empty(1)
    fn, ctx, pr, po, nd = g._lib.mcalf_loglike_batch, g._ctx, row.ctypes.data, out.ctypes.data, g.ndim
    t0 = time.perf_counter()
    for _ in range(n): fn(ctx, pr, 1, nd, 0, None, po, None)
    t_c = (time.perf_counter() - t0) / n
    g.reset_stats(); fn(ctx, pr, 1, nd, 0, None, po, None); kms = g.stats()["last_kernel_ms"] * 1e3
    print("cfg %d: lnlhood_worker %.1f us | raw C-ABI call %.1f us | kernel (events) %.1f us" % (cfg, t_py * 1e6, t_c * 1e6, kms))
