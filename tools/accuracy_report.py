"""Accuracy of the fp32 and fp64 kernels against the oracle over prior draws (GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, quick_bench as qb
from oracle import mcalf_oracle as orc
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
out = {}
for cfg, B in ((1, 512), (2, 384), (3, 96), (4, 96)):
    spec, kw = orc.config_kwargs(cfg, GOLD)
    o = orc.OracleFitter(spec, **kw)
    g = qb.make(cfg)
    U = np.random.default_rng(900 + cfg).random((B, o.ndim))
    P = np.array([o._scale_cube_pc(u) for u in U])
    ref = np.array([o.lnlhood_worker(p) for p in P])
    got = g.lnlhood_batch(P); got64 = g.lnlhood_batch(P, fp64=True)
    nf = min(B, 48)
    f32 = g.reconstruct_spec_batch(P[:nf]); f64 = g.reconstruct_spec_batch(P[:nf], fp64=True)
    fe32 = fe64 = 0.0
    for i in range(nf):
        m = o.reconstruct_spec(P[i]); c = abs(o.unpack(P[i])[1])
        fe32 = max(fe32, np.abs(f32[i] - m).max() / c); fe64 = max(fe64, np.abs(f64[i] - m).max() / c)
    rel32 = np.abs(got - ref) / np.abs(ref); rel64 = np.abs(got64 - ref) / np.abs(ref)
    out["cfg%d" % cfg] = dict(samples=B, logl_rel_fp32_max=float(rel32.max()), logl_rel_fp32_median=float(np.median(rel32)),
                              logl_rel_fp64_max=float(rel64.max()), flux_err_fp32_max=float(fe32), flux_err_fp64_max=float(fe64),
                              logl_range=[float(ref.min()), float(ref.max())])
    print("cfg", cfg, out["cfg%d" % cfg])
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(GOLD), "..", "gpurun_out", "accuracy.json")
json.dump(out, open(dst, "w"), indent=1)
