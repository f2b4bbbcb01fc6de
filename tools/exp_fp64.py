import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, numpy as np, quick_bench as qb
for cfg, B in ((4, 4096), (2, 16384)):
    g = qb.make(cfg)
    U = torch.rand((B, g.ndim), dtype=torch.float64, device='cuda')
    g.lnlhood_batch(U, unit_cube=True, fp64=True); torch.cuda.synchronize()
    t0 = time.perf_counter(); g.lnlhood_batch(U, unit_cube=True, fp64=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('cfg', cfg, 'fp64 kernel: %.1f ms for %d samples = %.0f logL/s' % (dt * 1e3, B, B / dt))
