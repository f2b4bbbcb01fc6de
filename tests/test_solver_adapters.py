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
