"""Development aid: kernel time of small device-resident batches (one CTA per sample, up to 32 warps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch, quick_bench as qb
for cfg in (4, 2):
    g = qb.make(cfg)
    for B in (1, 8, 37, 148, 200, 592, 1024):
        U = torch.rand((B, g.ndim), dtype=torch.float64, device='cuda')
        for thr in (0, 256, 512):
            g.set_option('threads', thr)
            ts = []
            for _ in range(20):
                g.lnlhood_batch(U, unit_cube=True); torch.cuda.synchronize()
                ts.append(g.stats()['last_kernel_ms'] * 1e3)
            print('cfg %d B %5d threads opt %4d -> %s : kernel %.1f us (min %.1f)' % (cfg, B, thr, g.geometry()['threads'], np.median(ts), min(ts)), flush=True)
        g.set_option('threads', 0)
