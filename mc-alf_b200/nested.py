"""A minimal batched nested sampler -- a reference consumer of the batched likelihood.

The reference hands its likelihood to third-party samplers (PolyChord, MultiNest, dynesty, jaxns; none
is installed in this image).  This module is NOT a replacement for them: it is the smallest sampler
that exercises the batched entry point the way a vectorised solver would -- every iteration proposes a
whole block of candidate points inside the bounding box of the live set (enlarged), evaluates them in
one kernel launch with the prior transform fused (``lnlhood_batch(U, unit_cube=True)``), and lets the
candidates, in draw order, replace the current worst live point whenever they beat it.  Uniform-in-box proposals are only efficient in few
dimensions (BASELINE config 1: ndim 4); it exists to demonstrate and test the glue, and writes the
reference's chain formats (``mcalf_b200.chains``).
"""
import numpy as np


def batched_nested_sampling(fitter, nlive=400, batch=2048, dlogz=0.05, max_iter=200000, enlarge=1.3, seed=0,
                            int_dims=None):
    """-> dict(logz, logz_err, samples [n, ndim] physical, logl [n], logw [n], ncall, nlaunch).

    ``int_dims``: unit-cube dimensions whose physical value is truncated to an integer (the ncomp slot,
    ``fitter.startind``): proposals there are drawn over the full unit interval.
    """
    rng = np.random.default_rng(seed)
    ndim = fitter.ndim
    int_dims = [fitter.startind] if int_dims is None else list(int_dims)
    U = rng.random((nlive, ndim))
    L = np.asarray(fitter.lnlhood_batch(U, unit_cube=True), dtype=float)
    ncall, nlaunch = nlive, 1
    logz, h, logx = -np.inf, 0.0, 0.0
    dead_u, dead_l, dead_logw = [], [], []
    log_shrink = -1.0 / nlive
    for it in range(max_iter):
        lo, hi = U.min(axis=0), U.max(axis=0)
        c, w = 0.5 * (lo + hi), 0.5 * (hi - lo) * enlarge
        lo, hi = np.clip(c - w, 0.0, 1.0), np.clip(c + w, 0.0, 1.0)
        lo[int_dims], hi[int_dims] = 0.0, 1.0
        cand = lo + (hi - lo) * rng.random((batch, ndim))
        cl = np.asarray(fitter.lnlhood_batch(cand, unit_cube=True), dtype=float)      # one launch per block
        ncall += batch
        nlaunch += 1
        used = 0
        # candidates in DRAW order (sorting them by likelihood would bias the shrinkage): one that beats the
        # current worst live point is a uniform draw from the constrained prior and replaces it
        for k in range(batch):
            worst = int(np.argmin(L))
            if not cl[k] > L[worst]:
                continue
            logw = logx + np.log1p(-np.exp(log_shrink)) + L[worst]
            dead_u.append(U[worst].copy()); dead_l.append(L[worst]); dead_logw.append(logw)
            logz = np.logaddexp(logz, logw)
            logx += log_shrink
            U[worst], L[worst] = cand[k], cl[k]
            used += 1
            if used >= nlive // 4:         # keep the box estimate fresh
                break
        if np.logaddexp(logz, logx + L.max()) - logz < dlogz:
            break
    # remaining live points
    logw_live = logx - np.log(nlive) + L
    for i in np.argsort(L):
        dead_u.append(U[i]); dead_l.append(L[i]); dead_logw.append(logw_live[i])
        logz = np.logaddexp(logz, logw_live[i])
    dead_u, dead_l, dead_logw = np.array(dead_u), np.array(dead_l), np.array(dead_logw)
    pw = np.exp(dead_logw - logz)
    info = float(np.sum(pw * (dead_l - logz)))
    samples = np.asarray(fitter.prior_transform_batch(dead_u))
    return dict(logz=float(logz), logz_err=float(np.sqrt(max(info, 0.0) / nlive)), samples=samples, logl=dead_l,
                logw=dead_logw - logz, ncall=ncall, nlaunch=nlaunch, iterations=it + 1)


def equal_weight_resample(result, n, seed=42):
    """Posterior draws with equal weights (what the reference writes to ``_equal_weights.txt``)."""
    p = np.exp(result["logw"])
    p /= p.sum()
    idx = np.random.default_rng(seed).choice(len(p), size=n, replace=True, p=p)
    return result["samples"][idx], result["logl"][idx]
