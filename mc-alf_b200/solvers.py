"""Solver-side glue for the batched entry points.

The nested samplers are third-party and absent from this image; these adapters only depend on the
small protocol each sampler uses, are import-guarded, and are tested against fakes
(tests/test_solver_adapters.py).

* ``BatchPool``        a ``pool``-like object for dynesty (``NestedSampler(..., pool=BatchPool(f),
                       queue_size=N)``; reference call site ``cli.py:196-206``): ``map(func, points)``
                       evaluates the whole list of proposals in ONE kernel launch when ``func`` is
                       the fitter's likelihood / prior transform, and falls back to the builtin
                       ``map`` for anything else dynesty maps over the pool.
* ``jax_likelihood``   the callable ``get_jax_likelihood()`` returns (``cli.py:237``): the CUDA path
                       behind ``jax.pure_callback`` with ``vmap_method="broadcast_all"``, so jaxns'
                       vmapped live-point block arrives as one batch.  jaxns draws the ncomp slot as
                       a continuous uniform and floors it (``cli.py:251``, ``hires_fitter.py:616``);
                       the kernel's ``int()`` of that slot does the same for non-negative values.
"""
import numpy as np


class BatchPool:
    def __init__(self, fitter, size=None):
        self.fitter = fitter
        self.size = size or 1
        self.launches = 0

    def _is(self, func, *names):
        target = getattr(func, "__func__", func)
        for name in names:
            bound = getattr(self.fitter, name)
            if target is getattr(bound, "__func__", bound) and getattr(func, "__self__", self.fitter) is self.fitter:
                return True
        # dynesty wraps callables (e.g. _function_wrapper with .func); unwrap one level
        inner = getattr(func, "func", None)
        return inner is not None and inner is not func and self._is(inner, *names)

    def map(self, func, iterable):
        pts = list(iterable)
        if not pts:
            return []
        if self._is(func, "lnlhood_dy", "lnlhood_worker"):
            self.launches += 1
            return list(self.fitter.lnlhood_batch(np.asarray(pts, dtype=np.float64)))
        if self._is(func, "lnlhood_pc"):
            self.launches += 1
            return [(v, []) for v in self.fitter.lnlhood_batch(np.asarray(pts, dtype=np.float64))]
        if self._is(func, "_scale_cube_pc"):
            self.launches += 1
            return list(self.fitter.prior_transform_batch(np.asarray(pts, dtype=np.float64)))
        return list(map(func, pts))

    # context-manager / lifecycle no-ops some samplers call on a pool
    def close(self):
        pass

    def join(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def jax_likelihood(fitter):
    """``p[ndim] (float32 or float64) -> logL`` usable inside jit/vmap (jaxns)."""
    import jax   # ImportError here is the right failure: jaxns needs jax
    import jax.numpy as jnp

    def host(p):
        p = np.asarray(p, dtype=np.float64)
        flat = p.reshape(-1, p.shape[-1])
        out = fitter.lnlhood_batch(flat)
        return out.reshape(p.shape[:-1]).astype(np.float32)

    def loglike(p):
        p = jnp.asarray(p)
        shape = jax.ShapeDtypeStruct(p.shape[:-1], jnp.float32)
        return jax.pure_callback(host, shape, p, vmap_method="broadcast_all")

    return loglike
