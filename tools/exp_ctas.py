import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, quick_bench as qb
for cfg, B in ((4, 32768), (3, 16384), (2, 65536)):
    g = qb.make(cfg)
    U = torch.rand((B, g.ndim), dtype=torch.float64, device='cuda')
    for dense in (0, 1):
        g.set_option('dense', dense)
        for thr in (128, 256):
            g.set_option('threads', thr)
            geo = g.geometry()
            ms = qb.timeit(g, U)
            print('cfg', cfg, 'dense', int(g.get_option('dense')), 'thr', thr, 'ctas', geo['ctas_per_sm'], 'smem', geo['smem_bytes'], '%.3f ms %.2f M/s' % (ms, B / ms / 1e3))
