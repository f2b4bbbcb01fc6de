"""Development aid: device-resident logL/s of the four BASELINE configs with the library the environment selects."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, quick_bench as qb
for cfg in (4, 2, 3, 1):
    g = qb.make(cfg)
    B = {1: 262144, 2: 131072, 3: 32768, 4: 131072}[cfg]
    U = torch.rand((B, g.ndim), dtype=torch.float64, device='cuda')
    ms = qb.timeit(g, U, reps=5)
    print('cfg %d  %.3f ms  %.2f M logL/s' % (cfg, ms, B / ms / 1e3), flush=True)
