"""TEST INFRASTRUCTURE: stand-ins with the public interface of the third-party samplers the drivers talk to
(dynesty, jax) -- none of them is installed in this image.  Only the calls ``mcalf_b200.drivers`` /
``mcalf_b200.solvers`` make are provided; the sampling itself is a crude importance sampler."""
import sys
import types

import numpy as np


class _Wrapper:                      # dynesty wraps user callables like this (``_function_wrapper``)
    def __init__(self, func, args=(), kwargs=None, name="f"):
        self.func, self.args, self.kwargs, self.name = func, args, kwargs or {}, name

    def __call__(self, x):
        return self.func(x, *self.args, **self.kwargs)


class _Results:
    pass


class DynamicNestedSampler:
    """Draws ``nblocks`` blocks of ``queue_size`` prior points, exactly the way dynesty uses a pool: the prior transform
    and the likelihood are mapped over whole blocks through ``pool.map``."""
    nblocks = 6

    def __init__(self, loglikelihood, prior_transform, ndim, bound='multi', method='auto', pool=None, queue_size=None,
                 rstate=None, **kw):
        assert bound == 'none' and method == 'unif'
        self.loglikelihood = _Wrapper(loglikelihood, name="loglikelihood")
        self.prior_transform = _Wrapper(prior_transform, name="prior_transform")
        self.ndim, self.pool, self.queue_size = ndim, pool, queue_size or 1
        self.rstate = rstate or np.random.default_rng(0)
        self.M = pool.map if pool is not None else map

    def run_nested(self, **kw):
        us, vs, ls = [], [], []
        for _ in range(self.nblocks):
            u = self.rstate.random((self.queue_size, self.ndim))
            v = np.array(list(self.M(self.prior_transform, list(u))))
            l = np.array(list(self.M(self.loglikelihood, list(v))), dtype=float)
            us.append(u); vs.append(v); ls.append(l)
        v, l = np.concatenate(vs), np.concatenate(ls)
        order = np.argsort(l)
        r = _Results()
        r.samples, r.logl = v[order], l[order]
        n = len(l)
        logw = r.logl - np.log(n)                                 # plain importance weights of prior draws
        r.logz = np.logaddexp.accumulate(logw)
        r.logzerr = np.full(n, 0.1)
        r.logwt = logw
        self.results = r


NestedSampler = DynamicNestedSampler


def resample_equal(samples, weights, rstate=None):
    rstate = rstate or np.random.default_rng(1)
    idx = rstate.choice(len(weights), size=len(weights), p=weights / weights.sum())
    return samples[idx]


def install_dynesty():
    mod = types.ModuleType("dynesty")
    utils = types.ModuleType("dynesty.utils")
    utils.resample_equal = resample_equal
    mod.DynamicNestedSampler, mod.NestedSampler, mod.utils = DynamicNestedSampler, NestedSampler, utils
    sys.modules["dynesty"], sys.modules["dynesty.utils"] = mod, utils
    return mod


def install_jax():
    """A ``jax`` whose ``pure_callback`` simply calls the host function and whose ``vmap`` hands the callback the whole
    batch at once -- the contract ``vmap_method='broadcast_all'`` gives the real one."""
    jax = types.ModuleType("jax")
    jnp = types.ModuleType("jax.numpy")
    jnp.asarray, jnp.float32 = np.asarray, np.float32

    class ShapeDtypeStruct:
        def __init__(self, shape, dtype):
            self.shape, self.dtype = tuple(shape), dtype

    calls = []

    def pure_callback(host, shape, *args, vmap_method=None):
        assert vmap_method == "broadcast_all"
        out = np.asarray(host(*args))
        calls.append(np.asarray(args[0]).shape)
        assert out.shape == shape.shape and out.dtype == shape.dtype
        return out

    def vmap(f):
        return lambda batch: f(np.asarray(batch))                # broadcast_all: one call with the leading axis kept

    jax.ShapeDtypeStruct, jax.pure_callback, jax.vmap, jax.numpy, jax.callback_shapes = ShapeDtypeStruct, pure_callback, vmap, jnp, calls
    sys.modules["jax"], sys.modules["jax.numpy"] = jax, jnp
    return jax


def uninstall(*names):
    for n in names:
        for k in [k for k in sys.modules if k == n or k.startswith(n + ".")]:
            del sys.modules[k]
