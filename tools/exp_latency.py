"""Latency / throughput vs batch size through the public API (host numpy path and device-tensor path)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, numpy as np, quick_bench as qb
for cfg in (2, 4):
    g = qb.make(cfg)
    rng = np.random.default_rng(0)
    for B in (1, 16, 200, 1024, 4096, 16384, 65536):
        U = rng.random((B, g.ndim)); Ud = torch.from_numpy(U).cuda()
        for _ in range(3): g.lnlhood_batch(U, unit_cube=True)
        n = max(3, min(200, 20000 // B))
        t0 = time.perf_counter()
        for _ in range(n): g.lnlhood_batch(U, unit_cube=True)
        th = (time.perf_counter() - t0) / n
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): g.lnlhood_batch(Ud, unit_cube=True)
        torch.cuda.synchronize(); td = (time.perf_counter() - t0) / n
        print('cfg %d B %6d  host path %9.1f us/call %10.0f logL/s | device path %9.1f us/call %10.0f logL/s' % (cfg, B, th * 1e6, B / th, td * 1e6, B / td))
    p = g._scale_cube_pc(rng.random(g.ndim))
    for _ in range(3): g.lnlhood_worker(p)
    t0 = time.perf_counter()
    for _ in range(300): g.lnlhood_worker(p)
    print('cfg %d scalar lnlhood_worker: %.1f us/call' % (cfg, (time.perf_counter() - t0) / 300 * 1e6))
