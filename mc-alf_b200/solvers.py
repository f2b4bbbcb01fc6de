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
import logging

import numpy as np


_UNWRAP_ATTRS = ("__wrapped__", "func", "loglikelihood", "prior_transform", "fn", "f")


def _unwrap_candidates(func, depth=4):
    """``func`` and whatever callables sampler-side wrappers hide it behind: dynesty's
    ``_function_wrapper`` (``.func``), ``functools.partial`` (``.func``/``.args[0]``), ``functools.wraps``
    (``__wrapped__``), objects carrying ``.loglikelihood`` / ``.prior_transform``."""
    seen, todo = [], [(func, 0)]
    while todo:
        f, d = todo.pop()
        if f is None or any(f is g for g in seen):
            continue
        seen.append(f)
        if d >= depth:
            continue
        for name in _UNWRAP_ATTRS:
            inner = getattr(f, name, None)
            if callable(inner):
                todo.append((inner, d + 1))
        args = getattr(f, "args", None)
        if isinstance(args, tuple) and args and callable(args[0]):
            todo.append((args[0], d + 1))
    return seen


class BatchPool:
    def __init__(self, fitter, size=None):
        self.fitter = fitter
        self.size = size or 1
        self.launches = 0
        self.scalar_fallbacks = 0
        self._warned = False

    def _is(self, func, *names):
        """True when ``func`` (possibly wrapped) is one of this fitter's bound methods ``names``."""
        targets = []
        for name in names:
            bound = getattr(self.fitter, name, None)
            if bound is not None:
                targets.append(getattr(bound, "__func__", bound))
        for f in _unwrap_candidates(func):
            if getattr(f, "__self__", None) is self.fitter and getattr(f, "__func__", f) in targets:
                return True
        return False

    def _owned(self, func):
        return any(getattr(f, "__self__", None) is self.fitter for f in _unwrap_candidates(func))

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
        if self._owned(func):
            # a method of OUR fitter that has no batched form here: correct, but one launch per point
            self.scalar_fallbacks += 1
            if not self._warned:
                self._warned = True
                logging.getLogger(__name__).warning(
                    "BatchPool: %r belongs to the fitter but has no batched form; mapping it point by point", func)
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
