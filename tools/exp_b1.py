import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, numpy as np, quick_bench as qb
for cfg in (2, 4):
    g = qb.make(cfg)
    rng = np.random.default_rng(0)
    for B in (1, 8, 64, 148, 592):
        U = rng.random((B, g.ndim)); Ud = torch.from_numpy(U).cuda()
        for thr in (0, 512, 1024):
            g.set_option('threads', thr)
            for _ in range(5): g.lnlhood_batch(U, unit_cube=True)
            n = 300
            t0 = time.perf_counter()
            for _ in range(n): g.lnlhood_batch(U, unit_cube=True)
            th = (time.perf_counter() - t0) / n
            kms = g.stats()['last_kernel_ms']
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(n): g.lnlhood_batch(Ud, unit_cube=True)
            torch.cuda.synchronize(); td = (time.perf_counter() - t0) / n
            print('cfg %d B %4d thr %4d: host %.1f us, device-path %.1f us, kernels %.1f us' % (cfg, B, g.geometry()['threads'], th * 1e6, td * 1e6, kms * 1e3))
