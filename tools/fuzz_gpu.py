"""Extended random-problem fuzz of the CUDA path against the oracle (GPU box; not part of the test suite)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mcalf_b200
from oracle import mcalf_oracle as orc
from tests.test_gpu_parity import _random_problem, const_term
worst_l, worst_f, n = 0.0, 0.0, 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    spec, kw = _random_problem(seed)
    o = orc.OracleFitter(spec, **kw)
    g = mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]),
                              **{k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items() if k not in ("fitrange", "fitlines", "ncomp")})
    U = np.random.default_rng(5000 + seed).random((32, o.ndim))
    P = np.array([o._scale_cube_pc(u) for u in U])
    with np.errstate(all="ignore"):
        ref = np.array([o.lnlhood_worker(p) for p in P])
    got = g.lnlhood_batch(P)
    C = abs(const_term(o))
    fin = np.isfinite(ref)
    assert np.array_equal(np.isinf(got), np.isinf(ref)), seed
    rel = np.abs(got[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), C)
    flux = g.reconstruct_spec_batch(P[:4])
    fe = max(np.abs(flux[i] - o.reconstruct_spec(P[i])).max() / abs(o.unpack(P[i])[1]) for i in range(4))
    worst_l, worst_f, n = max(worst_l, rel.max() if fin.any() else 0), max(worst_f, fe), n + 1
    if (fin.any() and rel.max() > 1e-6) or fe > 1e-6:
        print("FAIL seed", seed, rel.max(), fe, kw)
    g.close()
print("problems", n, "worst logL rel %.2e worst flux %.2e" % (worst_l, worst_f))
