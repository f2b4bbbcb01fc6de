"""Development aid: rate at one CTA size.  usage: exp_one.py <cfg> <threads>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, quick_bench as qb
cfg, thr = int(sys.argv[1]), int(sys.argv[2])
g = qb.make(cfg)
B = {1: 262144, 2: 131072, 3: 32768, 4: 131072}[cfg]
U = torch.rand((B, g.ndim), dtype=torch.float64, device='cuda')
g.set_option('threads', thr)
geo = g.geometry()
ms = qb.timeit(g, U, reps=5)
print('cfg %d thr %d ctas %d: %.3f ms %.2f M/s' % (cfg, geo['threads'], geo['ctas_per_sm'], ms, B / ms / 1e3))
