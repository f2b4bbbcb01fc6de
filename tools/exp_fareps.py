"""Development aid: rate and accuracy (against the fp64 kernel) as a function of the far-field tolerance."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch, quick_bench as qb
for cfg in (4, 3, 2):
    g = qb.make(cfg)
    B = {2: 131072, 3: 32768, 4: 131072}[cfg]
    U = torch.rand((B, g.ndim), dtype=torch.float64, device='cuda')
    Us = U[:512].cpu().numpy()
    ref = g.lnlhood_batch(Us, unit_cube=True, fp64=True)
    fref = g.reconstruct_spec_batch(Us[:64], unit_cube=True, fp64=True)
    for eps in (1e-9, 3e-9, 1e-8, 3e-8):
        g.set_option('far_eps', eps)
        ms = qb.timeit(g, U, reps=5)
        got = g.lnlhood_batch(Us, unit_cube=True)
        fl = g.reconstruct_spec_batch(Us[:64], unit_cube=True)
        g.set_option('collect_stats', 1); g.reset_stats(); g.lnlhood_batch(U[:4096], unit_cube=True); st = g.stats(); g.set_option('collect_stats', 0)
        print('cfg %d far_eps %.0e: %.2f M/s  far %.3f  max dlogL/logL %.2e  max flux err %.2e' % (
            cfg, eps, B / ms / 1e3, st['evals_far'] / st['evals_total'], np.abs(got / ref - 1).max(), np.abs(fl - fref).max()), flush=True)
