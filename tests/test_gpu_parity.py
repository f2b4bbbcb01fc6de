"""GPU: the CUDA path (through the C-ABI, via the als_fitter mirror) against the oracle and against
golden outputs of the unmodified reference.  Tolerances are BASELINE.json's: model flux within 1e-6
of the continuum, |dlogL| <= 1e-6 |logL| (with an absolute floor of 1e-6 |C|, C the constant term,
where logL crosses zero -- SURVEY.md section 7 hard part 3)."""
import numpy as np
import pytest

from oracle import mcalf_oracle as orc
from tests.cases import ALL_TAGS, case

pytestmark = pytest.mark.gpu

FLUX_TOL = 1e-6
LOGL_RTOL = 1e-6


def fitters(tag, **extra_gpu):
    import mcalf_b200
    spec, kw, extra = case(tag)
    o = orc.OracleFitter(spec, **kw, **extra)
    g = mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]),
                              **{k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()
                                 if k not in ("fitrange", "fitlines", "ncomp")}, **extra, **extra_gpu)
    return o, g


def logl_close(got, ref, const):
    got, ref = np.asarray(got), np.asarray(ref)
    assert np.array_equal(np.isinf(got), np.isinf(ref)), (got, ref)
    fin = np.isfinite(ref)
    tol = LOGL_RTOL * np.maximum(np.abs(ref[fin]), abs(const))
    err = np.abs(got[fin] - ref[fin])
    assert (err <= tol).all(), "max |dlogL|/tol = %g" % (err / tol).max()
    return (err / np.maximum(np.abs(ref[fin]), 1e-300)).max() if fin.any() else 0.0


def const_term(o):
    with np.errstate(all="ignore"):
        w = 1.0 / o.obj_noise ** 2
        return -0.5 * np.nansum(np.where(np.isnan(o.obj), np.nan, -np.log(w) + np.log(2 * np.pi)))


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("tag", ALL_TAGS)
def test_golden_reference_outputs(tag, precision, golden):
    """logL, chi2 and model flux computed by the UNMODIFIED reference (tests/golden)."""
    o, g = fitters(tag, precision=precision)
    P = golden[tag + "_P"]
    assert g.velstep == pytest.approx(float(golden[tag + "_velstep"]), rel=1e-14)
    assert np.array_equal(np.column_stack([g._blo, g._bhi]), golden[tag + "_bounds"])
    logl, chi2 = g.lnlhood_batch(P, return_chi2=True)
    logl_close(logl, golden[tag + "_logL"], const_term(o))
    ref_chi2 = golden[tag + "_chi2"]
    assert np.allclose(chi2, ref_chi2, rtol=2e-6 if precision == "fp32" else 1e-10)
    flux = golden[tag + "_flux"]
    got = g.reconstruct_spec_batch(P[:flux.shape[0]])
    cont = np.array([o.unpack(p)[1] for p in P[:flux.shape[0]]])[:, None]
    err = np.abs(got - flux) / np.abs(cont)
    assert err.max() <= (FLUX_TOL if precision == "fp32" else 1e-11), err.max()
    # the scalar callbacks are batches of one
    assert g.lnlhood_worker(P[0]) == logl[0] or (np.isinf(logl[0]) and np.isinf(g.lnlhood_worker(P[0])))
    assert g.lnlhood_pc(P[0])[1] == []
    assert np.array_equal(g.reconstruct_spec(P[0]), got[0])


def test_known_answers_testdata():
    """the reference's two mock spectra at their truth parameters."""
    o, g = fitters("cfg1_truth")
    assert g.lnlhood_worker(np.array([1.0, 13.8, 3.0, 15.0])) == pytest.approx(5001.865105876514, rel=1e-6)
    assert g.chi2(np.array([1.0, 13.8, 3.0, 15.0])) == pytest.approx(1956.6353392519727, rel=2e-6)
    o, g = fitters("cfg2_truth")
    p = np.array([10, 13.6, 2.999, 17.5, 13.0, 2.9995, 8, 13.8, 3.0, 20, 13.6, 3.001, 25, 13.2, 3.0005, 15, 13.4, 3.0015, 30,
                  13.5, 3.002, 10, 14.0, 3.0025, 25, 14.2, 3.0035, 15, 13.7, 3.0039, 20], dtype=float)
    assert g.lnlhood_worker(p) == pytest.approx(4991.860095571161, rel=1e-6)
    assert g.chi2(p) == pytest.approx(1976.6453598626786, rel=2e-6)


def test_mock_spectrum_is_model_plus_noise():
    """The reference's own golden vector through the CUDA model: Flux - N(0,.02;seed 42) == model(truth)."""
    o, g = fitters("cfg1_truth")
    np.random.seed(42)
    noise = np.random.normal(0, 0.02, size=len(g.obj_wl))
    m = g.reconstruct_spec(np.array([1.0, 13.8, 3.0, 15.0]))
    assert np.abs(g.obj - noise - m).max() < FLUX_TOL
    m64 = g.reconstruct_spec_batch(np.array([[1.0, 13.8, 3.0, 15.0]]), fp64=True)[0]
    assert np.abs(g.obj - noise - m64).max() < 1e-11   # the reference's own u formula is noisy at 1e-12


@pytest.mark.parametrize("cfg,B", [(1, 512), (2, 256), (3, 24), (4, 32)])
def test_prior_draws_vs_oracle(cfg, B):
    """Fresh prior draws (unit cube on the GPU, _scale_cube_pc in the oracle) at every BASELINE config."""
    o, g = fitters("cfg%d" % cfg)
    U = np.random.default_rng(100 + cfg).random((B, o.ndim))
    P = np.array([o._scale_cube_pc(u) for u in U])
    assert np.array_equal(g.prior_transform_batch(U), P)            # bit-exact prior transform
    assert np.array_equal(np.array([g._scale_cube_pc(u) for u in U]), P)
    Pmn = np.array([o._scale_cube_mn(u.copy()) for u in U])          # MultiNest form: no int() on the ncomp slot
    assert np.array_equal(g.prior_transform_batch(U, no_trunc=True), Pmn)
    assert np.array_equal(g.lnlhood_batch(U[:16], unit_cube=True, no_trunc=True), g.lnlhood_batch(Pmn[:16]))
    ref = np.array([o.lnlhood_worker(p) for p in P])
    got = g.lnlhood_batch(U, unit_cube=True)
    assert np.array_equal(got, g.lnlhood_batch(P))                  # same kernel either way
    worst = logl_close(got, ref, const_term(o))
    got64 = g.lnlhood_batch(P, fp64=True)
    assert np.allclose(got64, ref, rtol=1e-10)
    nf = min(B, 8)
    flux = g.reconstruct_spec_batch(P[:nf])
    for i in range(nf):
        m = o.reconstruct_spec(P[i])
        assert np.abs(flux[i] - m).max() / abs(o.unpack(P[i])[1]) <= FLUX_TOL
    print("cfg", cfg, "worst relative logL error", worst)


def test_onecomp_and_targonly():
    o, g = fitters("edge_gap")
    m = g.reconstruct_onecomp(10.0, 0.97, 14.0, 2.99502, 12.0)
    assert np.abs(m - o.reconstruct_onecomp(10.0, 0.97, 14.0, 2.99502, 12.0)).max() <= FLUX_TOL
    m = g.reconstruct_onecomp_fill(10.0, 1.0, 15.0, 23.801, 1.5)
    assert np.abs(m - o.reconstruct_onecomp(10.0, 1.0, 15.0, 23.801, 1.5, fill=True)).max() <= FLUX_TOL
    # an LSF wider than the halo sized from the bounds (specres 45 vs 10): re-routed, still right
    m = g.reconstruct_onecomp(45.0, 1.0, 14.0, 2.99502, 12.0)
    assert np.abs(m - o.reconstruct_onecomp(45.0, 1.0, 14.0, 2.99502, 12.0)).max() <= 1e-9
    p = np.array([2, 14.0, 2.99502, 12.0, 13.7, 2.99840, 20.0, 15.0, 23.7605, 9.0])
    assert np.abs(g.reconstruct_spec(p, targonly=True) - o.reconstruct_spec(p, targonly=True)).max() <= FLUX_TOL
    assert np.abs(g.reconstruct_spec(p) - o.reconstruct_spec(p)).max() <= FLUX_TOL


def test_large_damping_routes_to_fp64():
    """a > a_max is outside the fp32 series: the sample must land in the fp64 kernel, not be wrong."""
    spec, kw, _ = case("cfg1")
    atomic = {"FAKE 1548": (1548.204, 0.1899, 2.643e11)}    # gamma x1000 -> a ~ 0.13-3
    import mcalf_b200
    o = orc.OracleFitter(spec, **dict(kw, fitlines=["FAKE 1548"]), atomic=atomic)
    g = mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], ["FAKE 1548"], list(kw["ncomp"]),
                              specres=kw["specres"], contval=kw["contval"], atomic=atomic)
    P = np.array([[1, 14.0, 3.0, 2.0], [1, 13.5, 3.001, 25.0], [1, 15.0, 2.9995, 1.0]])
    g.reset_stats()
    got = g.lnlhood_batch(P)
    assert g.stats()["samples_fp64"] == 0 and g.stats()["kernel_launches"] == 2   # re-routed, not requested
    ref = np.array([o.lnlhood_worker(p) for p in P])
    assert np.allclose(got, ref, rtol=1e-9)
    flux = g.reconstruct_spec_batch(P)
    for i in range(3):
        assert np.abs(flux[i] - o.reconstruct_spec(P[i])).max() < 1e-9


def test_device_tensor_path_matches_host_path():
    import torch
    o, g = fitters("cfg2")
    U = np.random.default_rng(5).random((300, o.ndim))
    host = g.lnlhood_batch(U, unit_cube=True)
    dev = g.lnlhood_batch(torch.from_numpy(U).cuda(), unit_cube=True)
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), host)
    g.set_option("slice", 64)           # several pipelined slices, ragged tail
    assert np.array_equal(g.lnlhood_batch(U, unit_cube=True), host)
    fh = g.reconstruct_spec_batch(U[:70], unit_cube=True, dtype=np.float32)
    fd = g.reconstruct_spec_batch(torch.from_numpy(U[:70]).cuda(), unit_cube=True, dtype=torch.float32)
    assert np.array_equal(fd.cpu().numpy(), fh)


def test_launch_geometry_does_not_change_results():
    """Per-sample results are independent of CTA size / CTAs per SM / batch composition (the
    multi-GPU requirement: bit-identical whichever shard a sample lands in)."""
    o, g = fitters("cfg3")
    U = np.random.default_rng(9).random((64, o.ndim))
    base = g.lnlhood_batch(U, unit_cube=True)
    for threads in (128, 512, 1024):
        g.set_option("threads", threads)
        assert np.array_equal(g.lnlhood_batch(U, unit_cube=True), base)
    g.set_option("threads", 0)
    assert np.array_equal(g.lnlhood_batch(U[::-1].copy(), unit_cube=True)[::-1], base)
    assert np.array_equal(g.lnlhood_batch(U[10:11], unit_cube=True), base[10:11])


def test_empty_and_bad_input():
    import mcalf_b200
    o, g = fitters("cfg1")
    assert g.lnlhood_batch(np.zeros((0, g.ndim))).shape == (0,)
    with pytest.raises(ValueError):
        g.lnlhood_batch(np.zeros((3, g.ndim - 1)))
    out = g.lnlhood_batch(np.array([[1.0, np.nan, 3.0, 15.0]]))
    assert np.isnan(out[0])
    with pytest.raises(mcalf_b200.capi.McalfError):
        g.set_option("no_such_option", 1)


def test_full_size_properties():
    """BASELINE cfg 4 at a bench-sized batch: properties that need no oracle run.
    (i) a sample's logL does not depend on where it sits in the batch; (ii) ncomp = 0 gives the
    continuum-only chi-square exactly; (iii) far-off-window components change nothing measurable;
    (iv) chi2 and logL are tied by logL = C - chi2/2."""
    o, g = fitters("cfg4")
    B = 16384
    U = np.random.default_rng(4).random((B, o.ndim))
    logl, chi2 = g.lnlhood_batch(U, unit_cube=True, return_chi2=True)
    assert np.isfinite(logl).all()
    C = const_term(o)
    assert np.allclose(logl, C - 0.5 * chi2, rtol=1e-14)
    idx = np.random.default_rng(0).permutation(B)
    assert np.array_equal(g.lnlhood_batch(U[idx], unit_cube=True), logl[idx])
    P = g.prior_transform_batch(U[:4])
    P[:, o.startind] = 0
    cont = P[:, 1]
    with np.errstate(all="ignore"):
        expect = np.array([np.nansum((o.obj - c) ** 2 / o.obj_noise ** 2) for c in cont])
    assert np.allclose(g.chi2_batch(P), expect, rtol=1e-6)
    # spot-check 4 samples of the big batch against the oracle
    Pq = g.prior_transform_batch(U[:4])
    ref = np.array([o.lnlhood_worker(p) for p in Pq])
    logl_close(logl[:4], ref, C)


def _random_problem(seed):
    """A small random problem: odd pixel counts, 1-3 windows (possibly with gaps), random line sets,
    free/fixed specres and continuum, fillers, NaN / zero-error pixels."""
    rng = np.random.default_rng(seed)
    names = list(orc.ATOMIC)
    lines = list(rng.choice(names, size=rng.integers(1, 5), replace=False))
    w0 = orc.ATOMIC[lines[0]][0]
    z0 = rng.uniform(1.5, 3.5)
    centre = w0 * (1 + z0)
    nwin = int(rng.integers(1, 4))
    velstep = rng.uniform(0.7, 3.0)
    waves, fitrange = [], []
    lam = centre * (1 - 400.0 / orc.C_KMS * rng.uniform(0.5, 1.5))
    for _ in range(nwin):
        n = int(rng.integers(3, 700))
        if rng.random() < 0.5:
            seg = lam * np.exp(np.arange(n) * velstep / orc.C_KMS)            # log-uniform
        else:
            seg = lam + np.arange(n) * lam * velstep / orc.C_KMS              # linear in wavelength
        waves.append(seg)
        fitrange.append((seg[0] - 1e-3, seg[-1] + 1e-3))
        lam = seg[-1] * (1 + rng.uniform(0.0005, 0.003))
    wave = np.concatenate(waves)
    flux = 1.0 + rng.normal(0, 0.03, wave.size)
    err = rng.uniform(0.01, 0.05, wave.size)
    if rng.random() < 0.5:
        flux[rng.integers(0, wave.size)] = np.nan
        err[rng.integers(0, wave.size)] = 0.0
    ncmax = int(rng.integers(0, 6))
    kw = dict(fitrange=fitrange, fitlines=lines, ncomp=(int(rng.integers(0, ncmax + 1)), ncmax), nfill=int(rng.integers(0, 3)),
              specres=[4.0, 14.0] if rng.random() < 0.6 else [float(rng.uniform(0.3, 12.0))],
              contval=[0.8, 1.2] if rng.random() < 0.5 else [1.0], Nrange=(11.5, 15.5), brange=(1.5, 45.0),
              zrange=(z0 - 0.002, z0 + 0.004))
    return (wave, flux, err), kw


@pytest.mark.parametrize("seed", range(12))
def test_random_small_problems(seed):
    import mcalf_b200
    spec, kw = _random_problem(seed)
    o = orc.OracleFitter(spec, **kw)
    g = mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]),
                              **{k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()
                                 if k not in ("fitrange", "fitlines", "ncomp")})
    assert g.velstep == pytest.approx(o.velstep, rel=1e-14)
    U = np.random.default_rng(1000 + seed).random((48, o.ndim))
    P = np.array([o._scale_cube_pc(u) for u in U])
    assert np.array_equal(g.prior_transform_batch(U), P)
    with np.errstate(all="ignore"):
        ref = np.array([o.lnlhood_worker(p) for p in P])
    got = g.lnlhood_batch(U, unit_cube=True)
    logl_close(got, ref, const_term(o))
    assert np.allclose(g.lnlhood_batch(P, fp64=True), ref, rtol=1e-9)
    flux = g.reconstruct_spec_batch(P[:6])
    for i in range(6):
        assert np.abs(flux[i] - o.reconstruct_spec(P[i])).max() / abs(o.unpack(P[i])[1]) <= FLUX_TOL


def test_long_spectrum_many_chunks():
    """More chunks than lanes (20000 px = 79 chunks) and more lines than one slot list row."""
    import mcalf_b200
    rng = np.random.default_rng(3)
    wave = 5000.0 * np.exp(np.arange(20001) * 1.3 / orc.C_KMS)
    spec = (wave, 1.0 + rng.normal(0, 0.02, wave.size), np.full(wave.size, 0.02))
    lines = ["CIV 1548", "CIV 1550", "SiIV 1393", "SiIV 1402"]
    kw = dict(fitrange=[(wave[0] - 1, wave[-1] + 1)], fitlines=lines, ncomp=(25, 25), nfill=1, specres=[5.0, 9.0],
              contval=[0.95, 1.05], Nrange=(12.0, 14.8), brange=(4.0, 35.0), zrange=(2.30, 2.46))
    o = orc.OracleFitter(spec, **kw)
    g = mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], lines, list(kw["ncomp"]),
                              **{k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()
                                 if k not in ("fitrange", "fitlines", "ncomp")})
    assert g.geometry()["nchunks"] > 64
    U = np.random.default_rng(8).random((6, o.ndim))
    P = np.array([o._scale_cube_pc(u) for u in U])
    ref = np.array([o.lnlhood_worker(p) for p in P])
    logl_close(g.lnlhood_batch(P), ref, const_term(o))
    flux = g.reconstruct_spec_batch(P[:2])
    for i in range(2):
        assert np.abs(flux[i] - o.reconstruct_spec(P[i])).max() / abs(o.unpack(P[i])[1]) <= FLUX_TOL


def test_problem_too_large_is_refused():
    import mcalf_b200
    wave = 4000.0 * np.exp(np.arange(120000) * 1.0 / orc.C_KMS)
    spec = (wave, np.ones(wave.size), np.full(wave.size, 0.02))
    with pytest.raises(mcalf_b200.capi.McalfError) as ei:
        mcalf_b200.als_fitter(spec, [[wave[0] - 1, wave[-1] + 1]], ["CIV 1548"], [1, 1])
    assert ei.value.code == mcalf_b200.capi.E_RESOURCE


def test_bench_size_batch_fp32_vs_fp64_kernel():
    """BASELINE cfg 4 at the bench's own batch size (262144 unit-cube vectors): the fp32 kernel against
    the fp64 check kernel (itself tied to the oracle at 1e-10 above) on 4096 randomly chosen samples of
    that batch, plus the size-independent properties on the whole batch."""
    o, g = fitters("cfg4")
    B = 262144
    U = np.random.default_rng(4000).random((B, o.ndim))
    logl, chi2 = g.lnlhood_batch(U, unit_cube=True, return_chi2=True)
    assert np.isfinite(logl).all()
    C = const_term(o)
    assert np.allclose(logl, C - 0.5 * chi2, rtol=1e-14)
    pick = np.random.default_rng(1).choice(B, size=4096, replace=False)
    ref = g.lnlhood_batch(U[pick], unit_cube=True, fp64=True)
    worst = logl_close(logl[pick], ref, C)
    assert np.array_equal(g.lnlhood_batch(U[pick], unit_cube=True), logl[pick])    # batch composition is irrelevant
    st = g.stats()
    assert st["samples_fp64"] == 4096            # nothing of the fp32 batch was silently re-routed
    # ... and 64 of them directly against the oracle (numpy + scipy wofz), not only against the repo's own fp64 kernel
    spot = pick[:64]
    Pq = g.prior_transform_batch(U[spot])
    ref_o = np.array([o.lnlhood_worker(p) for p in Pq])
    worst_o = logl_close(logl[spot], ref_o, C)
    assert np.allclose(ref[:64], ref_o, rtol=1e-10)
    print("bench-size batch: worst relative logL difference fp32 kernel vs oracle %.2e (64 rows)" % worst_o)
    print("bench-size batch: worst relative logL difference fp32 vs fp64 kernel %.2e" % worst)


def test_damped_and_saturated_lines():
    """Column densities far above the BASELINE priors (sub-DLA / DLA Lyman-alpha, logN 17-21): damping
    wings fill the window, tau reaches 1e7 at the core; nothing may take a form it is not valid for."""
    import mcalf_b200
    rng = np.random.default_rng(11)
    wave = 4700.0 * np.exp(np.arange(6000) * 2.0 / orc.C_KMS)
    spec = (wave, 1.0 + rng.normal(0, 0.02, wave.size), np.full(wave.size, 0.02))
    kw = dict(fitrange=[(wave[0] - 1, wave[-1] + 1)], fitlines=["HI 1215", "HI 1025"], ncomp=(3, 3), nfill=0,
              specres=[6.0, 12.0], contval=[1.0], Nrange=(12.0, 21.0), brange=(5.0, 60.0), zrange=(2.92, 2.98))
    o = orc.OracleFitter(spec, **kw)
    g = mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]),
                              **{k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()
                                 if k not in ("fitrange", "fitlines", "ncomp")})
    P = np.array([[8.0, 3, 20.3, 2.95, 30.0, 17.2, 2.93, 12.0, 13.5, 2.97, 8.0],
                  [7.0, 3, 21.0, 2.94, 55.0, 19.0, 2.96, 20.0, 18.0, 2.925, 5.0],
                  [11.0, 3, 18.5, 2.95, 40.0, 16.0, 2.951, 7.0, 14.2, 2.9495, 25.0],
                  [9.0, 2, 19.7, 2.975, 15.0, 12.3, 2.93, 33.0, 20.0, 2.95, 10.0]])
    ref = np.array([o.lnlhood_worker(p) for p in P])
    logl_close(g.lnlhood_batch(P), ref, const_term(o))
    assert np.allclose(g.lnlhood_batch(P, fp64=True), ref, rtol=1e-9)
    flux = g.reconstruct_spec_batch(P)
    for i in range(len(P)):
        assert np.abs(flux[i] - o.reconstruct_spec(P[i])).max() <= FLUX_TOL
    U = np.random.default_rng(12).random((64, o.ndim))
    Pd = np.array([o._scale_cube_pc(u) for u in U])
    logl_close(g.lnlhood_batch(Pd), np.array([o.lnlhood_worker(p) for p in Pd]), const_term(o))


def test_multi_device_fitter_single_process():
    """Several context replicas driven from host threads of one process (all visible GPUs; two replicas
    on the same GPU when there is only one) return exactly what a single context returns."""
    import torch
    from mcalf_b200.multi import MultiDeviceFitter
    spec, kw, _ = case("cfg2")
    o, g = fitters("cfg2")
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0]
    m = MultiDeviceFitter(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]), devices=devices,
                          **{k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()
                             if k not in ("fitrange", "fitlines", "ncomp")})
    U = np.random.default_rng(21).random((1001, o.ndim))
    assert np.array_equal(m.lnlhood_batch(U, unit_cube=True), g.lnlhood_batch(U, unit_cube=True))
    assert np.array_equal(m.prior_transform_batch(U[:7]), g.prior_transform_batch(U[:7]))
    assert np.array_equal(m.reconstruct_spec_batch(U[:5], unit_cube=True), g.reconstruct_spec_batch(U[:5], unit_cube=True))
    assert m.ndim == g.ndim and m.lnlhood_dy(g._scale_cube_pc(U[0])) == g.lnlhood_dy(g._scale_cube_pc(U[0]))
    assert m.lnlhood_batch(U[:1], unit_cube=True).shape == (1,)
    m.close()


def test_batched_nested_sampling_recovers_the_mock_truth(tmp_path):
    """BASELINE config 1 end to end with the built-in batched sampler: the single-component CIV mock
    (truth N 13.8, z 3.0, b 15.0) fitted with whole blocks of proposals per launch, chains written in the
    reference's formats and re-evaluated in one batch."""
    import time
    from mcalf_b200 import chains
    from mcalf_b200.nested import batched_nested_sampling, equal_weight_resample
    o, g = fitters("cfg1")
    t0 = time.perf_counter()
    r = batched_nested_sampling(g, nlive=400, batch=4096, dlogz=0.1, seed=1)
    dt = time.perf_counter() - t0
    samples, logl = equal_weight_resample(r, 1000)
    med = np.median(samples, axis=0)
    assert med[0] == 1.0
    assert abs(med[1] - 13.8) < 0.03 and abs(med[2] - 3.0) < 2e-5 and abs(med[3] - 15.0) < 1.0, med
    assert logl.max() > 4995.0                      # truth logL 5001.87
    base = str(tmp_path / "chain_0")
    chains.write_stats(base, r["logz"], r["logz_err"])
    chains.write_equal_weights(base, logl, samples)
    stored, again = chains.logl_of_chain(g, base)
    assert np.allclose(stored, again, rtol=1e-12)
    print("nested sampling: %d likelihood calls in %d launches, %.2f s, logZ %.2f +/- %.2f, median %s"
          % (r["ncall"], r["nlaunch"], dt, r["logz"], r["logz_err"], med))


@pytest.mark.parametrize("a0", [1e-5, 1e-4, 1e-3, 1e-2])
def test_device_faddeeva_vs_wofz(a0):
    """Re w(u + i a) as the kernels compute it (mcalf_voigt_h) against scipy.special.wofz -- the
    reference's own Faddeeva (hires_fitter.py:365): fp32 forms to 5e-7 relative, fp64 form to 1e-12."""
    from scipy.special import wofz
    from mcalf_b200 import capi
    rng = np.random.default_rng(5)
    u = np.concatenate([rng.uniform(-12, 12, 200000), rng.uniform(-3000, 3000, 50000)]).astype(np.float32).astype(float)
    a = np.full_like(u, np.float32(a0))
    ref = wofz(u + 1j * a).real
    got32 = capi.voigt_h(u, a, mode=0)
    assert (np.abs(got32 - ref) / ref).max() < 5e-7
    got64 = capi.voigt_h(u, a, mode=1)
    assert (np.abs(got64 - ref) / ref).max() < 1e-12
    big = capi.voigt_h(rng.uniform(-30, 30, 20000), np.full(20000, 2.5), mode=1)      # large damping: fp64 only
    assert np.isfinite(big).all()


@pytest.mark.parametrize("a0", [1e-5, 1e-4, 1e-3, 1e-2])
def test_device_faddeeva_short_core_form(a0):
    """The short line-core form used for weak lines (kappa <= 8; one-float u, MUFU.EX2): absolute error
    below 2.5e-7 and relative error below 2e-6 everywhere inside the core, which bounds the flux error
    F * kappa * dH of a kappa = 8 line by 1.5e-7."""
    from scipy.special import wofz
    from mcalf_b200 import capi
    u = np.random.default_rng(6).uniform(-6, 6, 300000).astype(np.float32).astype(float)
    a = np.full_like(u, np.float32(a0))
    ref = wofz(u + 1j * a).real
    got = capi.voigt_h(u, a, mode=2)
    assert np.abs(got - ref).max() < 2.5e-7
    assert (np.abs(got - ref) / ref).max() < 2e-6
    kappa = 8.0
    assert (kappa * np.exp(-kappa * ref) * np.abs(got - ref)).max() < 1.5e-7     # measured 1.2e-7; the bar is 1e-6


def test_derived_quantities():
    """Equivalent width and total column density against the oracle restatement."""
    o, g = fitters("cfg2")
    P = np.array([o._scale_cube_pc(u) for u in np.random.default_rng(31).random((12, o.ndim))])
    for lineid in (0, 1):
        ref = np.array([o.calc_w(p, lineid) for p in P])
        got = g.calc_w_batch(P, lineid)
        assert np.allclose(got, ref, rtol=2e-6, atol=1e-7), np.abs(got / ref - 1).max()
    assert g.calc_w(P[0]) == pytest.approx(o.calc_w(P[0]), rel=2e-6)
    assert g.calc_N(P[0]) == pytest.approx(o.calc_N(P[0]), rel=1e-14)
