"""Development aid: rate of the fp32 kernel against the CTA geometry (threads per CTA, CTAs per SM)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, quick_bench as qb
cfgs = [int(a) for a in sys.argv[1:]] or [4]
for cfg in cfgs:
    g = qb.make(cfg)
    B = {1: 262144, 2: 131072, 3: 32768, 4: 65536}[cfg]
    U = torch.rand((B, g.ndim), dtype=torch.float64, device='cuda')
    for thr in (0, 128, 160, 192, 224, 256, 288, 320, 384, 512):
        g.set_option('ctas_per_sm', 0)
        try:
            g.set_option('threads', thr)
        except Exception as e:
            print('cfg', cfg, 'thr', thr, e); continue
        occ = g.geometry()['ctas_per_sm']
        for ctas in sorted({occ, max(occ - 1, 1)}, reverse=True):
            g.set_option('ctas_per_sm', ctas)
            geo = g.geometry()
            ms = qb.timeit(g, U, reps=4)
            print('cfg %d thr %4d ctas %d warps/SM %2d smem %6d  %.3f ms %.2f M/s' % (cfg, geo['threads'], geo['ctas_per_sm'], geo['threads'] // 32 * geo['ctas_per_sm'], geo['smem_bytes'], ms, B / ms / 1e3), flush=True)
