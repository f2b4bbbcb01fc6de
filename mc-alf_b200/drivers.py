"""Solver drivers: the hand-off of the likelihood to the nested samplers, with the batched path.

The reference's ``cli.py`` builds each sampler inline (PolyChord ``:82-120``, dyPolyChord ``:122-160``,
MultiNest ``:161-188``, dynesty ``:190-206``, jaxns ``:208-326``).  PolyChord / MultiNest call the likelihood
one point at a time from Fortran and need nothing new (``als_fitter.lnlhood_pc`` / ``lnlhood_mn`` are drop-in).
dynesty and jaxns can hand over whole blocks of points; these two drivers are that glue -- what
BASELINE.json's north star calls the "batched vectorised path for jaxns/dynesty" -- written against the
samplers' public interfaces and free of the defects the shipped branches have (``comp``
undefined and ``dyfunc`` never imported in the dynesty branch, whose output also lacks the two leading
columns ``pc_analyzer`` expects).  The samplers are third-party and absent from this image: imports are
guarded, and the tests drive the code through stand-in modules with the same interface.

Every driver writes the reference's chain formats (``mcalf_b200.chains``) under ``filesbasename``.
"""
import time

import numpy as np

from . import chains
from .solvers import BatchPool


def _finish(fitter, filesbasename, logz, logz_err, samples, logl=None):
    """Equal-weight samples -> ``.stats`` + ``_equal_weights.txt`` (cli.py:292-326); logL re-evaluated in ONE launch
    when the sampler did not hand it back."""
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    if logl is None:
        logl = np.asarray(fitter.lnlhood_batch(samples))
    chains.write_stats(filesbasename, logz, logz_err)
    chains.write_equal_weights(filesbasename, logl, samples)
    return np.asarray(logl)


def run_dynesty(fitter, filesbasename, queue_size=256, dynamic=True, sampler_kwargs=None, run_kwargs=None, seed=None):
    """dynesty with every iteration's proposals evaluated in one kernel launch (reference call site
    ``cli.py:196-197``: ``DynamicNestedSampler(temp.lnlhood_dy, temp._scale_cube_pc, temp.ndim, bound='none',
    method='unif')``).  The sampler gets ``pool=BatchPool(fitter)`` and ``queue_size``: dynesty then maps the prior
    transform and the likelihood over blocks of ``queue_size`` points, which the pool turns into batched calls."""
    try:
        import dynesty
        from dynesty import utils as dyfunc
    except ImportError as e:
        raise ImportError("dynesty is required for the dynesty solver. Please install it.") from e
    pool = BatchPool(fitter, size=queue_size)
    kw = dict(bound='none', method='unif')                       # the reference's choices
    kw.update(sampler_kwargs or {})
    if seed is not None:
        kw.setdefault("rstate", np.random.default_rng(seed))
    cls = dynesty.DynamicNestedSampler if dynamic else dynesty.NestedSampler
    sampler = cls(fitter.lnlhood_dy, fitter._scale_cube_pc, fitter.ndim, pool=pool, queue_size=queue_size, **kw)
    t0 = time.perf_counter()
    sampler.run_nested(**(run_kwargs or {}))
    elapsed = time.perf_counter() - t0
    res = sampler.results
    logz = float(res.logz[-1])
    logz_err = float(res.logzerr[-1]) if getattr(res, "logzerr", None) is not None else float("nan")
    weights = np.exp(np.asarray(res.logwt) - logz)               # normalised weights (cli.py:201)
    weights /= weights.sum()
    samples_equal = dyfunc.resample_equal(np.asarray(res.samples), weights)
    logl = _finish(fitter, filesbasename, logz, logz_err, samples_equal)
    return dict(logz=logz, logz_err=logz_err, samples=samples_equal, logl=logl, seconds=elapsed,
                launches=pool.launches, scalar_fallbacks=pool.scalar_fallbacks, results=res)


def run_jaxns(fitter, filesbasename, max_samples=1e5, num_live_points=500, difficult_model=False, device="gpu", seed=43):
    """jaxns with the CUDA likelihood behind ``jax.pure_callback`` (reference ``cli.py:208-326``): jaxns vmaps the
    likelihood over its live points, so every block arrives at ``als_fitter.lnlhood_batch`` as one batch."""
    try:
        import jax
        import jax.numpy as jnp
        from jaxns import Model, NestedSampler, Prior
        from jaxns.utils import resample
        from tensorflow_probability.substrates import jax as tfp
    except ImportError as e:
        raise ImportError("jaxns is required for the jaxns solver. Please install it.") from e
    if device == "cpu":
        jax.config.update("jax_platform_name", "cpu")
    log_likelihood = fitter.get_jax_likelihood()
    lowers = jnp.asarray(np.array([np.min(b) for b in fitter.bounds]), dtype=jnp.float32)
    uppers = jnp.asarray(np.array([np.max(b) for b in fitter.bounds]), dtype=jnp.float32)

    def prior_model():
        p = yield Prior(tfp.distributions.Uniform(low=lowers, high=uppers), name='p')
        return p

    model = Model(prior_model=prior_model, log_likelihood=log_likelihood)
    ns = NestedSampler(model=model, max_samples=int(max_samples), num_live_points=int(num_live_points),
                       difficult_model=bool(difficult_model))
    t0 = time.perf_counter()
    termination_reason, state = ns(key=jax.random.PRNGKey(int(seed)))
    results = ns.to_results(state=state, termination_reason=termination_reason)
    elapsed = time.perf_counter() - t0
    key = jax.random.PRNGKey(42)
    raw = results.samples['p'] if 'p' in results.samples else list(results.samples.values())[0]
    S = int(max_samples)
    samples_equal = np.asarray(resample(key=key, samples=raw, log_weights=results.log_dp_mean, S=S, replace=True))
    logl_equal = np.asarray(resample(key=key, samples=results.log_L_samples, log_weights=results.log_dp_mean, S=S, replace=True))
    samples_equal = samples_equal.reshape(samples_equal.shape[0], -1)
    _finish(fitter, filesbasename, float(results.log_Z_mean), float(results.log_Z_uncert), samples_equal, logl_equal.reshape(-1))
    return dict(logz=float(results.log_Z_mean), logz_err=float(results.log_Z_uncert), samples=samples_equal,
                logl=logl_equal.reshape(-1), seconds=elapsed, results=results)


def run_batched(fitter, filesbasename, nlive=400, batch=4096, dlogz=0.05, nequal=1000, seed=0):
    """The package's own minimal batched nested sampler (``mcalf_b200.nested``): no third-party solver needed;
    efficient only in few dimensions (uniform-in-box proposals)."""
    from .nested import batched_nested_sampling, equal_weight_resample
    t0 = time.perf_counter()
    r = batched_nested_sampling(fitter, nlive=nlive, batch=batch, dlogz=dlogz, seed=seed)
    elapsed = time.perf_counter() - t0
    samples, logl = equal_weight_resample(r, nequal)
    _finish(fitter, filesbasename, r["logz"], r["logz_err"], samples, logl)
    return dict(logz=r["logz"], logz_err=r["logz_err"], samples=samples, logl=logl, seconds=elapsed,
                launches=r["nlaunch"], ncall=r["ncall"])
