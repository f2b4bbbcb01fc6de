"""Development aid: the 48-register build forced on / off at every config."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, quick_bench as qb
for cfg in (4, 3, 2, 1):
    g = qb.make(cfg)
    B = {1: 262144, 2: 131072, 3: 32768, 4: 131072}[cfg]
    U = torch.rand((B, g.ndim), dtype=torch.float64, device='cuda')
    for dense in (0, 1):
        g.set_option('dense', dense)
        geo = g.geometry()
        ms = qb.timeit(g, U, reps=5)
        print('cfg %d dense %d (in use %d) thr %d ctas %d: %.3f ms %.2f M/s' % (cfg, dense, g.get_option('dense'), geo['threads'], geo['ctas_per_sm'], ms, B / ms / 1e3), flush=True)
