"""GPU: parity hardening beyond tests/test_gpu_parity.py -- the calling conventions, launch geometries, host-layer
paths and batch compositions that round-1's review found untested.  Everything goes
through the C-ABI; the oracle (numpy + scipy wofz) is the checker."""
import ctypes
import threading

import numpy as np
import pytest

from oracle import mcalf_oracle as orc
from tests.cases import case
from tests.test_gpu_parity import FLUX_TOL, _random_problem, const_term, fitters, logl_close

pytestmark = pytest.mark.gpu


def _gpu(spec, kw, **extra):
    import mcalf_b200
    return mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]),
                                 **{k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()
                                    if k not in ("fitrange", "fitlines", "ncomp")}, **extra)


def test_lnlhood_mn_multinest_calling_convention():
    """MultiNest hands the likelihood a ctypes double* plus (ndim, nparam) and the prior transform works in
    place on that buffer (hires_fitter.py:211-216, 274-285)."""
    o, g = fitters("cfg2")
    U = np.random.default_rng(17).random((6, o.ndim))
    for u in U:
        buf = (ctypes.c_double * (o.ndim + 3))(*u, 0.0, 0.0, 0.0)            # nparam > ndim: extra slots are ignored
        cube = ctypes.cast(buf, ctypes.POINTER(ctypes.c_double))
        ret = g._scale_cube_mn(cube, o.ndim, o.ndim + 3)
        assert ret is cube
        ref_p = o._scale_cube_mn(u.copy())
        assert np.array_equal(np.array([cube[i] for i in range(o.ndim)]), ref_p)
        got = g.lnlhood_mn(cube, o.ndim, o.ndim + 3)
        assert isinstance(got, float)
        ref = o.lnlhood_worker(ref_p)
        logl_close([got], [ref], const_term(o))
        assert got == g.lnlhood_batch(ref_p[None, :])[0]                     # the scalar call is a batch of one


@pytest.mark.parametrize("cfg,B", [(3, 128), (4, 128)])
def test_many_prior_draws_logl_and_flux_vs_oracle(cfg, B):
    """128 fresh prior draws at the two large BASELINE configs, logL AND model flux of every draw."""
    o, g = fitters("cfg%d" % cfg)
    U = np.random.default_rng(500 + cfg).random((B, o.ndim))
    P = g.prior_transform_batch(U)
    flux = g.reconstruct_spec_batch(P)
    logl = g.lnlhood_batch(P)
    ref_logl = np.empty(B)
    worst = 0.0
    w = 1.0 / o.obj_noise ** 2
    for i, p in enumerate(P):
        m = o.reconstruct_spec(p)
        worst = max(worst, np.abs(flux[i] - m).max() / abs(o.unpack(p)[1]))
        ref_logl[i] = -0.5 * np.nansum(w * (o.obj - m) ** 2 - np.log(w) + np.log(2.0 * np.pi))   # :292-294
    assert worst <= FLUX_TOL, worst
    rel = logl_close(logl, ref_logl, const_term(o))
    print("cfg %d: %d draws, worst flux error %.2e, worst relative logL error %.2e" % (cfg, B, worst, rel))


@pytest.mark.parametrize("seed", range(12, 112))
def test_random_problems_fuzz(seed):
    """100 further random problems (odd sizes, 1-3 windows with gaps, random line sets, fillers, NaN and
    zero-error pixels): logL and the flux of every draw against the oracle."""
    spec, kw = _random_problem(seed)
    o = orc.OracleFitter(spec, **kw)
    g = _gpu(spec, kw)
    U = np.random.default_rng(2000 + seed).random((24, o.ndim))
    P = np.array([o._scale_cube_pc(u) for u in U])
    with np.errstate(all="ignore"):
        ref = np.array([o.lnlhood_worker(p) for p in P])
        got = g.lnlhood_batch(U, unit_cube=True)
        logl_close(got, ref, const_term(o))
        flux = g.reconstruct_spec_batch(P)
        for i in range(len(P)):
            assert np.abs(flux[i] - o.reconstruct_spec(P[i])).max() / abs(o.unpack(P[i])[1]) <= FLUX_TOL
    g.close()


def test_asymmlike_veto_fires_for_some_rows_only():
    """Asymmlike (hires_fitter.py:296-303): -inf when too many pixels sit > 4/5 sigma ABOVE the model.  A batch
    that mixes good fits (no veto) with models far below the data (veto) must match the oracle row by row."""
    spec, kw, _ = case("cfg1")
    kw = dict(kw, Asymmlike=True, contval=[0.5, 1.2])
    extra = {"gauss_cdf": (3, 1, 0)}
    o = orc.OracleFitter(spec, **kw, **extra)
    g = _gpu(spec, kw, **extra)
    rng = np.random.default_rng(3)
    P = np.array([[c, 1, 13.8 + rng.normal(0, 0.05), 3.0 + rng.normal(0, 2e-5), 15.0 + rng.normal(0, 1)]
                  for c in np.concatenate([np.linspace(0.9, 1.06, 24), np.full(8, 1.0)])])
    ref = np.array([o.lnlhood_worker(p) for p in P])
    got = g.lnlhood_batch(P)
    assert np.isneginf(ref).sum() >= 8 and np.isfinite(ref).sum() >= 8       # the veto splits the batch
    logl_close(got, ref, const_term(o))
    got64 = g.lnlhood_batch(P, fp64=True)
    assert np.array_equal(np.isinf(got64), np.isinf(ref)) and np.allclose(got64[np.isfinite(ref)], ref[np.isfinite(ref)], rtol=1e-10)


def test_rows_longer_than_the_cta():
    """ndim larger than the CTA: a short window (few chunks -> 64-thread CTAs) with 30 components,
    ndim 93, and explicitly threads = 32."""
    rng = np.random.default_rng(5)
    wave = 6180.0 * np.exp(np.arange(700) * 1.1 / orc.C_KMS)
    spec = (wave, 1.0 + rng.normal(0, 0.02, wave.size), np.full(wave.size, 0.02))
    kw = dict(fitrange=[(wave[0] - 0.1, wave[-1] + 0.1)], fitlines=["CIV 1548", "CIV 1550"], ncomp=(25, 30), nfill=0,
              specres=[6.0, 9.0], contval=[0.95, 1.05], Nrange=(12.0, 14.0), brange=(5.0, 30.0))
    o = orc.OracleFitter(spec, **kw)
    g = _gpu(spec, kw)
    assert o.ndim == 93 and g.geometry()["threads"] < o.ndim
    U = rng.random((40, o.ndim))
    P = np.array([o._scale_cube_pc(u) for u in U])
    ref = np.array([o.lnlhood_worker(p) for p in P])
    base = g.lnlhood_batch(U, unit_cube=True)
    logl_close(base, ref, const_term(o))
    assert np.allclose(g.lnlhood_batch(P, fp64=True), ref, rtol=1e-10)
    for threads in (32, 64, 256):
        g.set_option("threads", threads)
        assert np.array_equal(g.lnlhood_batch(U, unit_cube=True), base)
        assert np.array_equal(g.lnlhood_batch(P[:3], fp64=True), g.lnlhood_batch(P[:3], fp64=True))
    assert np.array_equal(g.prior_transform_batch(U), P)


def test_contexts_of_different_size_share_a_device():
    """a small context created after a big one must not lower the big one's shared-memory limit
    (the attribute is per function and device).  Big (8192 px) -> small -> big again, on the paths that
    launch the fp64 kernel with its opt-in shared memory."""
    o4, g4 = fitters("cfg4")
    U = np.random.default_rng(1).random((3000, o4.ndim))
    before = g4.lnlhood_batch(U, unit_cube=True)                             # > 1024 rows: pipelined path + fp64 fix-up launch
    f64 = g4.lnlhood_batch(U[:8], unit_cube=True, fp64=True)
    o1, g1 = fitters("cfg1")
    small = g1.lnlhood_batch(np.random.default_rng(2).random((3000, o1.ndim)), unit_cube=True)
    assert np.isfinite(small).all()
    assert np.array_equal(g4.lnlhood_batch(U, unit_cube=True), before)
    assert np.array_equal(g4.lnlhood_batch(U[:8], unit_cube=True, fp64=True), f64)
    w = g4.calc_w_batch(g4.prior_transform_batch(U[:4]))                      # single-line model in the same context
    assert np.isfinite(w).all() and np.array_equal(g4.lnlhood_batch(U, unit_cube=True), before)


def test_device_calls_on_different_streams():
    """Successive MCALF_F_ON_DEVICE calls of one context on different torch streams share the context's work
    counters: the library orders them; every result must equal the single-stream one."""
    import torch
    o, g = fitters("cfg2")
    U = torch.from_numpy(np.random.default_rng(9).random((4096, o.ndim))).cuda()
    ref = g.lnlhood_batch(U, unit_cube=True).clone()
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(4)]
    outs = []
    for rep in range(3):
        for st in streams:
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                outs.append(g.lnlhood_batch(U, unit_cube=True))
    torch.cuda.synchronize()
    for out in outs:
        assert torch.equal(out, ref)


def test_concurrent_use_of_one_context_is_refused_not_corrupted():
    """One call at a time per context: a second thread entering while a batch is in flight gets
    MCALF_E_INVALID ('context busy') -- or, if the first call already finished, a correct result."""
    import mcalf_b200
    o, g = fitters("cfg4")
    U = np.random.default_rng(2).random((65536, o.ndim))
    small = U[:16]
    expect_small = g.lnlhood_batch(small, unit_cube=True)
    outcome = {}

    def big():
        outcome["big"] = g.lnlhood_batch(U, unit_cube=True)

    busy = 0
    for attempt in range(4):
        th = threading.Thread(target=big)
        th.start()
        for _ in range(50):
            try:
                got = g.lnlhood_batch(small, unit_cube=True)
                assert np.array_equal(got, expect_small)
            except mcalf_b200.capi.McalfError as e:
                assert e.code == mcalf_b200.capi.E_INVALID and "busy" in str(e)
                busy += 1
        th.join()
        assert np.array_equal(outcome["big"][:16], expect_small)
    print("busy refusals seen:", busy)


def test_batch_pool_drives_the_real_fitter():
    """The dynesty pool adapter (cli.py:196-206) against the real fitter: one launch per mapped block,
    bit-identical to lnlhood_batch / prior_transform_batch, wrapped callables included."""
    import functools
    from mcalf_b200.solvers import BatchPool
    o, g = fitters("cfg2")
    rng = np.random.default_rng(4)
    U = rng.random((200, o.ndim))
    P = g.prior_transform_batch(U)
    pool = BatchPool(g)
    g.reset_stats()
    got = pool.map(g.lnlhood_dy, list(P))
    assert g.stats()["kernel_launches"] <= 2 and pool.launches == 1
    assert np.array_equal(np.array(got), g.lnlhood_batch(P))
    assert [v for v, _ in pool.map(g.lnlhood_pc, list(P))] == got
    assert np.array_equal(np.array(pool.map(g._scale_cube_pc, list(U))), P)

    class Wrapped:                              # dynesty's _function_wrapper shape
        def __init__(self, func):
            self.func, self.args, self.kwargs = func, (), {}

        def __call__(self, x):
            return self.func(x)

    assert pool.map(Wrapped(g.lnlhood_dy), list(P)) == got
    assert pool.map(functools.partial(g.lnlhood_dy), list(P)) == got
    assert pool.launches == 5 and pool.scalar_fallbacks == 0
    # scalar callbacks agree with the batch bit for bit
    assert [g.lnlhood_dy(p) for p in P[:5]] == got[:5]


def test_oneline_model_and_weak_forms_on_device():
    """MCALF_F_ONELINE (one line of one component) against the oracle's single-line model, and the weak-line
    forms (wide-interval wing polynomial beyond s = 16, short core form inside) against wofz."""
    from scipy.special import wofz
    from mcalf_b200 import capi
    o, g = fitters("cfg3")
    rows = np.array([[0.0, 1.0, 13.9, 2.9995, 17.0, li] for li in range(o.numlines)])
    flux = g.reconstruct_oneline_batch(rows)
    for li in range(o.numlines):
        ref = o._transmission(13.9, 2.9995, 17.0, o.linepars[li])
        assert np.abs(flux[li] - ref).max() <= FLUX_TOL
    u = np.random.default_rng(6).uniform(-9, 9, 200000).astype(np.float32).astype(float)
    for a0 in (1e-4, 6e-4):
        a = np.full_like(u, np.float32(a0))
        ref = wofz(u + 1j * a).real
        got = capi.voigt_h(u, a, mode=3)
        s = u * u + a * a
        assert (np.abs(got - ref) <= 1.2e-7 * np.exp(-np.minimum(s, 16.0) + 16.0) * (s >= 16) + 2.5e-7 * (s < 16) + 6e-6 * ref).all()


def test_dynesty_driver_and_jax_adapter_drive_the_real_fitter(tmp_path):
    """The batched solver glue (drivers.run_dynesty: cli.py:190-206 repaired; get_jax_likelihood: cli.py:237) against the
    REAL fitter, the samplers replaced by stand-ins with their public interface: every block of points is one launch,
    and what comes back equals lnlhood_batch."""
    from tests import fake_solvers
    from mcalf_b200 import chains, drivers
    o, g = fitters("cfg1")
    fake_solvers.install_dynesty()
    try:
        base = str(tmp_path / "dy_0")
        g.reset_stats()
        out = drivers.run_dynesty(g, base, queue_size=128, seed=3)
        nb = fake_solvers.DynamicNestedSampler.nblocks
        assert out["launches"] == 2 * nb and out["scalar_fallbacks"] == 0
        assert g.stats()["kernel_launches"] <= 2 * nb + 2 + 2          # prior kernels + likelihood kernels (+ fix-up) + the final re-evaluation
        stored, again = chains.logl_of_chain(g, base)
        assert np.allclose(stored, again, rtol=1e-12)
        ref = np.array([o.lnlhood_worker(p) for p in out["samples"][:16]])
        logl_close(out["logl"][:16], ref, const_term(o))
    finally:
        fake_solvers.uninstall("dynesty")
    jax = fake_solvers.install_jax()
    try:
        ll = g.get_jax_likelihood()
        P = g.prior_transform_batch(np.random.default_rng(8).random((200, o.ndim))).astype(np.float32)
        got = jax.vmap(ll)(P)
        assert got.dtype == np.float32 and jax.callback_shapes[-1] == (200, o.ndim)     # one callback for the whole block
        assert np.array_equal(got, g.lnlhood_batch(P.astype(np.float64)).astype(np.float32))
    finally:
        fake_solvers.uninstall("jax")


def test_register_lean_build_gives_the_same_bits():
    """The 48-register instantiation (five CTAs per SM, chosen automatically on long spectra) is the same arithmetic:
    forcing it on or off must not change a bit."""
    o, g = fitters("cfg4")
    U = np.random.default_rng(12).random((3000, o.ndim))
    auto = g.lnlhood_batch(U, unit_cube=True)
    assert g.get_option("dense") == 1.0 and g.geometry()["ctas_per_sm"] == 5      # the automatic choice at cfg 4 on a B200
    g.set_option("dense", 0)
    assert g.get_option("dense") == 0.0 and g.geometry()["ctas_per_sm"] == 4
    assert np.array_equal(g.lnlhood_batch(U, unit_cube=True), auto)
    g.set_option("dense", 1)
    assert np.array_equal(g.lnlhood_batch(U, unit_cube=True), auto)
    o2, g2 = fitters("cfg2")
    assert g2.get_option("dense") == 0.0                                            # short spectrum: not chosen
