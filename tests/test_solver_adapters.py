"""CPU: solver adapters against fakes (the samplers themselves are not installed)."""
import numpy as np

from mcalf_b200.solvers import BatchPool


class FakeFitter:
    """Stands in for als_fitter: same method names, counts batched launches."""
    ndim = 3

    def __init__(self):
        self.batches = []

    def lnlhood_batch(self, P):
        self.batches.append(len(P))
        return -np.sum(np.asarray(P) ** 2, axis=1)

    def prior_transform_batch(self, U):
        self.batches.append(len(U))
        return np.asarray(U) * 2.0

    def lnlhood_worker(self, p):
        return float(self.lnlhood_batch(np.asarray(p)[None, :])[0])

    def lnlhood_dy(self, p):
        return self.lnlhood_worker(p)

    def lnlhood_pc(self, p):
        return self.lnlhood_worker(p), []

    def _scale_cube_pc(self, u):
        return np.asarray(u) * 2.0


class Wrapper:           # how dynesty wraps user callables
    def __init__(self, func):
        self.func = func

    def __call__(self, x):
        return self.func(x)


def test_pool_batches_likelihood_calls():
    f = FakeFitter()
    pool = BatchPool(f)
    pts = [np.array([1.0, 2.0, 3.0]) * k for k in range(5)]
    out = pool.map(f.lnlhood_dy, pts)
    assert f.batches == [5] and pool.launches == 1
    assert np.allclose(out, [-14.0 * k * k for k in range(5)])
    out = pool.map(Wrapper(f.lnlhood_dy), pts)          # wrapped callable is recognised too
    assert f.batches == [5, 5]
    out = pool.map(f.lnlhood_pc, pts)
    assert out[2] == (-56.0, []) and f.batches == [5, 5, 5]
    out = pool.map(f._scale_cube_pc, pts)
    assert np.array_equal(out[1], pts[1] * 2) and f.batches[-1] == 5
    assert pool.map(f.lnlhood_dy, []) == []


def test_pool_unwraps_nested_wrappers():
    import functools
    f = FakeFitter()
    pool = BatchPool(f)
    pts = [np.ones(3) * k for k in range(4)]

    class DynestyLike:                      # dynesty's _function_wrapper keeps the callable in .func plus args/kwargs
        def __init__(self, func):
            self.func, self.args, self.kwargs = func, (), {}

        def __call__(self, x):
            return self.func(x, *self.args, **self.kwargs)

    @functools.wraps(f.lnlhood_dy)
    def wrapped(x):
        return f.lnlhood_dy(x)

    for fn in (DynestyLike(Wrapper(f.lnlhood_dy)), functools.partial(f.lnlhood_dy), wrapped, DynestyLike(f._scale_cube_pc)):
        f.batches.clear()
        pool.map(fn, pts)
        assert f.batches == [4], fn           # one batched launch, not four scalar calls
    # a method of our fitter that has no batched form: served point by point, and counted
    f.batches.clear()
    pool.map(f.lnlhood_worker.__self__.prior_transform_batch, [pts[0][None, :]])
    assert pool.scalar_fallbacks == 1


def test_pool_falls_back_for_foreign_functions():
    f = FakeFitter()
    pool = BatchPool(f)
    assert pool.map(lambda x: x + 1, [1, 2, 3]) == [2, 3, 4]
    assert f.batches == []
    other = FakeFitter()
    pool.map(other.lnlhood_dy, [np.zeros(3)])            # another fitter's method: not batched through ours
    assert f.batches == [] and other.batches == [1]


def test_batched_nested_sampler_on_a_gaussian():
    """The built-in batched sampler against a likelihood with a known evidence (CPU, fake fitter)."""
    from mcalf_b200.nested import batched_nested_sampling, equal_weight_resample

    class Gauss:
        ndim, startind = 3, 0          # dimension 0 plays the (ignored) integer slot
        mu, sig = np.array([0.5, 0.4, 0.6]), 0.05

        def lnlhood_batch(self, U, unit_cube=False):
            d = (np.asarray(U)[:, 1:] - self.mu[1:]) / self.sig
            return -0.5 * np.sum(d * d, axis=1)

        def prior_transform_batch(self, U):
            return np.asarray(U)

    r = batched_nested_sampling(Gauss(), nlive=300, batch=1024, seed=3)
    expect = 2 * np.log(0.05 * np.sqrt(2 * np.pi))          # two Gaussian dimensions well inside the unit cube
    assert abs(r["logz"] - expect) < 0.25, (r["logz"], expect)
    s, _ = equal_weight_resample(r, 2000)
    assert np.allclose(s[:, 1:].mean(axis=0), [0.4, 0.6], atol=0.01)
    assert r["nlaunch"] * 1024 + 300 >= r["ncall"]


def test_dynesty_driver_batches_every_block(tmp_path):
    """``drivers.run_dynesty`` against a stand-in dynesty: every block of proposals reaches the fitter as ONE batched
    call (prior transform and likelihood), the chain files come out in the reference's formats, and the reference's
    own reader understands them."""
    from tests import fake_solvers
    from mcalf_b200 import chains, drivers
    fake_solvers.install_dynesty()
    try:
        f = FakeFitter()
        f.bounds = [(0.0, 2.0)] * 3
        base = str(tmp_path / "chain_0")
        out = drivers.run_dynesty(f, base, queue_size=64, seed=5)
        nb = fake_solvers.DynamicNestedSampler.nblocks
        assert f.batches[:2 * nb] == [64] * (2 * nb)             # nblocks x (prior transform, likelihood), 64 points each
        assert out["launches"] == 2 * nb and out["scalar_fallbacks"] == 0
        assert f.batches[-1] == len(out["samples"])              # the equal-weight samples re-evaluated in one launch
        lnz, err, lh, post = chains.read_chains(base, return_sorted=False)
        assert lnz == out["logz"] and np.allclose(lh, out["logl"]) and post.shape == out["samples"].shape
        first = open(base + "_equal_weights.txt").readline().split()
        assert len(first) == 2 + 3 and float(first[0]) == 1.0 and float(first[1]) == -2.0 * out["logl"][0]
    finally:
        fake_solvers.uninstall("dynesty")


def test_jax_likelihood_adapter_batches_under_vmap():
    """``solvers.jax_likelihood`` (what ``get_jax_likelihood()`` returns, cli.py:237) against a stand-in jax: a vmapped
    block of live points reaches ``lnlhood_batch`` once, float32 in and out as jaxns runs it."""
    from tests import fake_solvers
    from mcalf_b200.solvers import jax_likelihood
    jax = fake_solvers.install_jax()
    try:
        f = FakeFitter()
        ll = jax_likelihood(f)
        one = ll(np.array([1.0, 2.0, 3.0], dtype=np.float32))
        assert one.shape == () and one.dtype == np.float32 and float(one) == -14.0 and f.batches == [1]
        block = np.arange(30, dtype=np.float32).reshape(10, 3)
        out = jax.vmap(ll)(block)
        assert out.shape == (10,) and out.dtype == np.float32 and f.batches == [1, 10]
        assert np.allclose(out, -np.sum(block.astype(np.float64) ** 2, axis=1))
    finally:
        fake_solvers.uninstall("jax")


def test_drivers_need_their_sampler():
    import pytest
    from mcalf_b200 import drivers
    with pytest.raises(ImportError, match="dynesty is required"):
        drivers.run_dynesty(FakeFitter(), "/tmp/none")
    with pytest.raises(ImportError, match="jaxns is required"):
        drivers.run_jaxns(FakeFitter(), "/tmp/none")
