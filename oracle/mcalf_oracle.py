"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MC-ALF likelihood hot path.

A NumPy/SciPy fp64 restatement of the reference's per-sample Voigt-model likelihood
(``mcalf/routines/hires_fitter.py``, reference tree at /root/reference).  Nothing in the product
package imports this file; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, and only as the checker / CPU baseline.

Parity status: PINNED for the CIV-doublet / single-window / 8 km/s case by the reference's own mock
spectra (``tests/golden/civ_mock*.npz``: ``Flux - N(0,0.02; seed 42)`` equals the model at the truth
parameters to 1e-15) and, for every other case exercised by the tests (multi-window, floating
specres/continuum, fillers, NaN pixels), by golden outputs of the UNMODIFIED reference run in the
build container through ``oracle/refshim.py`` (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
The reference has no tests of its own; third-party arithmetic is ``scipy.special.wofz`` (unpinned in
the reference's pyproject.toml:15-22; scipy 1.18.1 here) and astropy's ``convolve`` (restated).

Each function cites the reference lines it follows.
"""
import math

import numpy as np
from scipy.special import wofz

C_KMS = 2.9979245e5    # hires_fitter.py:65
C_CGS = 2.9979245e10   # hires_fitter.py:66
TAU_CONST = 0.014971475  # hires_fitter.py:364  (sqrt(pi) e^2 / (m_e c))
FWHM_TO_SIGMA = 2.354820  # hires_fitter.py:454
TRUNC_SIGMAS = 3.0348     # hires_fitter.py:458
FILLER_WREST = 250.0      # hires_fitter.py:121

# wrest [Angstrom], f, gamma [1/s] -- same table as oracle/refshim.py (CIV rows pinned, others unverified)
ATOMIC = {
    "CIV 1548": (1548.204, 0.1899, 2.643e8),
    "CIV 1550": (1550.781, 0.09475, 2.628e8),
    "HI 1215": (1215.67, 0.4164, 6.265e8),
    "HI 1025": (1025.7222, 0.07912, 1.897e8),
    "HI 972": (972.5367, 0.0290, 8.127e7),
    "SiIV 1393": (1393.7602, 0.513, 8.80e8),
    "SiIV 1402": (1402.7729, 0.254, 8.62e8),
}
# hires_fitter.py:101-110: CrII overrides of (f, gamma)
CRII_OVERRIDES = {"CrII 2066": (0.0512, 4.17e8), "CrII 2062": (0.0759, 4.06e8), "CrII 2056": (0.103, 4.07e8)}


def read_spectrum(specfile, coldef=("Wave", "Flux", "Err")):
    """astropy.io.ascii.read of a '# Wave Flux Err' table (hires_fitter.py:69-72)."""
    with open(specfile) as fh:
        names = fh.readline().lstrip("#").split()
    data = np.loadtxt(specfile, ndmin=2)
    cols = {n: data[:, i] for i, n in enumerate(names)}
    return tuple(np.asarray(cols[c], dtype=float) for c in coldef)


def sigma_clipped_median(values, sigma=3.0, maxiters=5):
    """Median after astropy-style sigma clipping; only the median is consumed (hires_fitter.py:84-87)."""
    d = np.asarray(values, dtype=float)
    for _ in range(maxiters):
        keep = np.abs(d - np.median(d)) <= sigma * d.std()
        if keep.all():
            break
        d = d[keep]
    return float(np.median(d))


def voigt_tau(wave_cm, logN, z, b_cms, wrest_cm, f, gamma):
    """Optical depth of one line (hires_fitter.py:331-367), cgs units."""
    cold = 10.0 ** logN
    nujk = C_CGS / wrest_cm
    dnu = b_cms / wrest_cm
    avoigt = gamma / (4.0 * np.pi * dnu)
    uvoigt = ((C_CGS / (wave_cm / (z + 1.0))) - nujk) / dnu
    cne = TAU_CONST * cold * f
    return cne * wofz(uvoigt + 1j * avoigt).real / dnu


def lsf_convolve_wrap(spec, fwhm_kms, velstep):
    """Periodic Gaussian LSF convolution (hires_fitter.py:452-464 + astropy semantics):
    sigma in pixels, half-width n = ceil(3.0348 sigma), taps exp(-k^2/2 sigma^2) renormalised by
    their truncated sum, array treated as periodic (boundary='wrap')."""
    sigma = (fwhm_kms / FWHM_TO_SIGMA) / velstep
    n = int(np.ceil(TRUNC_SIGMAS * sigma))
    k = np.arange(-n, n + 1)
    taps = np.exp(-0.5 * (k / sigma) ** 2)
    taps /= taps.sum()
    spec = np.asarray(spec, dtype=float)
    idx = (np.arange(spec.size)[:, None] + k[None, :]) % spec.size
    return (spec[idx] * taps[None, :]).sum(axis=1)


class OracleFitter:
    """Restatement of ``als_fitter`` state + likelihood (hires_fitter.py:32-200, 202-216, 236-248,
    287-328, 409-464).  ``spectrum`` is a path or a (wave, flux, err) tuple of arrays."""

    def __init__(self, spectrum, fitrange, fitlines, ncomp, nfill=0, specres=(7.0,), contval=(1.0,),
                 Nrange=(11.5, 16), brange=(1, 30), zrange=None, Nrangefill=(11.5, 16),
                 brangefill=(1, 30), wrangefill=None, coldef=("Wave", "Flux", "Err"),
                 Asymmlike=False, gauss_cdf=None, atomic=None):
        if isinstance(spectrum, (str, bytes)):
            wl, fl, er = read_spectrum(spectrum, coldef)
        else:
            wl, fl, er = (np.asarray(a, dtype=float) for a in spectrum)
        self.specres = list(specres)
        self.contval = list(contval)
        self.ncompmin, self.ncompmax = int(ncomp[0]), int(ncomp[1])
        self.nfill = int(nfill)
        self.freecont = len(self.contval) > 1          # :54-57
        self.freespecres = len(self.specres) > 1       # :59-62
        self.Asymmlike = bool(Asymmlike)

        ok = np.zeros(wl.shape, dtype=bool)            # :75-82 strict window mask
        for lo, hi in fitrange:
            ok |= (wl > lo) & (wl < hi)
        self.obj, self.obj_noise, self.obj_wl = fl[ok], er[ok], wl[ok]
        self.fitrange = [tuple(r) for r in fitrange]
        self.numfitranges = len(self.fitrange)
        steps = (self.obj_wl[1:] - self.obj_wl[:-1]) / self.obj_wl[1:] * C_KMS   # :84
        self.velstep = sigma_clipped_median(steps)                                # :85-87

        table = dict(ATOMIC)
        if atomic:
            table.update(atomic)
        self.fitlines = list(fitlines)
        self.numlines = len(self.fitlines)
        self.linepars = []
        for name in self.fitlines:                     # :93-113
            if name not in table:
                raise KeyError("line %r not in atomic table" % name)
            wrest, f, gamma = table[name]
            if name in CRII_OVERRIDES:
                f, gamma = CRII_OVERRIDES[name]
            self.linepars.append((float(wrest), float(f), float(gamma)))
        w0, f0, g0 = self.linepars[0]
        self.linefill = (FILLER_WREST, f0, g0)         # :120-121

        z_lims = []                                    # :134-149
        for k in range(self.ncompmax):
            if zrange is None:
                zlo = (self.fitrange[0][0] + 0.25) / w0 - 1.0
                zhi = (self.fitrange[0][1] - 0.25) / w0 - 1.0
            elif len(zrange) == 2:
                zlo, zhi = zrange
            elif len(zrange) >= 2 * self.ncompmax:
                zlo, zhi = zrange[2 * k], zrange[2 * k + 1]
            else:
                raise ValueError("Zrange keyword not understood")
            z_lims.append((zlo, zhi))
        z_lims_fill = []                               # :152-166
        for k in range(self.nfill):
            if wrangefill is None:
                lo, hi = np.min(self.obj_wl) + 0.25, np.max(self.obj_wl) - 0.25
            elif len(wrangefill) == 2:
                lo, hi = wrangefill
            elif len(wrangefill) == 2 * self.nfill:
                lo, hi = wrangefill[2 * k], wrangefill[2 * k + 1]
            else:
                raise ValueError("Wrangefill keyword not understood")
            z_lims_fill.append((lo / FILLER_WREST - 1.0, hi / FILLER_WREST - 1.0))

        self.startind = int(self.freecont) + int(self.freespecres)      # :169-174
        self.endind = self.startind + 3 * self.ncompmax + 1             # :176
        self.bounds = []                                                # :184-198
        if self.freespecres:
            self.bounds.append(tuple(self.specres))
        if self.freecont:
            self.bounds.append(tuple(self.contval))
        self.bounds.append(tuple(ncomp))
        for k in range(self.ncompmax):
            self.bounds += [tuple(Nrange), tuple(z_lims[k]), tuple(brange)]
        for k in range(self.nfill):
            self.bounds += [tuple(Nrangefill), tuple(z_lims_fill[k]), tuple(brangefill)]
        self.ndim = len(self.bounds)                                    # :200
        # :179-181 -- the reference draws these from an UNSEEDED normal; tests inject them
        self.gauss_cdf = list(gauss_cdf) if gauss_cdf is not None else [0, 0, 0]
        self.gracenum = 0.01 * len(self.obj)

    # ---- prior transforms (hires_fitter.py:202-216) ----
    def _scale_cube_pc(self, cube):
        out = np.array(cube, dtype=float)
        for i in range(len(out)):
            out[i] = out[i] * np.ptp(self.bounds[i]) + np.min(self.bounds[i])
            if i == self.startind:
                out[i] = int(out[i])
        return out

    def _scale_cube_mn(self, cube, ndim=None, nparam=None):
        n = self.ndim if ndim is None else ndim
        for i in range(n):
            cube[i] = cube[i] * np.ptp(self.bounds[i]) + np.min(self.bounds[i])
        return cube

    # ---- model synthesis (hires_fitter.py:369-377, 409-449) ----
    def _transmission(self, logN, z, b, line):
        wrest, f, gamma = line
        tau = voigt_tau(self.obj_wl / 1e8, logN, z, b * 1e5, wrest / 1e8, f, gamma)
        return np.exp(-1.0 * tau)

    def unpack(self, p):
        """(specres, continuum, thisncomp) exactly as reconstruct_spec parses p (:412-428)."""
        if self.freespecres:
            res = p[0]
        else:
            res = float(max(self.specres))
        if self.freecont:
            cont = p[1] if self.freespecres else p[0]
        else:
            cont = self.contval[0]
        return res, cont, int(p[self.startind])

    def reconstruct_spec(self, p, targonly=False):
        res, cont, thisncomp = self.unpack(p)
        model = np.ones_like(self.obj)
        s = self.startind
        for comp in range(thisncomp):
            logN, z, b = p[1 + 3 * comp + s:4 + 3 * comp + s]
            for line in self.linepars:
                model *= self._transmission(logN, z, b, line)
        if not targonly:
            for k in range(self.nfill):
                logN, z, b = p[3 * k + self.endind:3 * k + 3 + self.endind]
                model *= self._transmission(logN, z, b, self.linefill)
        if res > self.velstep:                          # :445
            model = lsf_convolve_wrap(model, res, self.velstep)
        return model * cont

    def reconstruct_onecomp(self, specresolution, continuum, N, z, b, fill=False):
        """hires_fitter.py:379-406"""
        model = np.ones_like(self.obj)
        for line in ([self.linefill] if fill else self.linepars):
            model *= self._transmission(N, z, b, line)
        if specresolution > self.velstep:
            model = lsf_convolve_wrap(model, specresolution, self.velstep)
        return model * continuum

    # ---- reductions (hires_fitter.py:236-248, 287-328) ----
    def chi2(self, p):
        model = self.reconstruct_spec(p)
        with np.errstate(all="ignore"):
            w = 1.0 / self.obj_noise ** 2
            return float(np.nansum(w * (self.obj - model) ** 2))

    def lnlhood_worker(self, p):
        model = self.reconstruct_spec(p)
        with np.errstate(all="ignore"):
            w = 1.0 / self.obj_noise ** 2
            lhood = -0.5 * np.nansum(w * (self.obj - model) ** 2 - np.log(w) + np.log(2.0 * np.pi))
            if self.Asymmlike:                          # :296-303
                resid = (self.obj - model) / self.obj_noise
                if (resid > 5).sum() > self.gauss_cdf[2] + self.gracenum:
                    return -np.inf
                if (resid > 4).sum() > self.gauss_cdf[1] + self.gracenum:
                    return -np.inf
        return float(lhood)

    # ---- derived quantities (hires_fitter.py:467-505), with the CURRENT parameter layout: the reference
    # indexes p[3*comp+startind] (no ncomp slot, :482/:499) and loops to ncompmax; the restatement follows
    # the layout reconstruct_spec uses (:431) and the active components only
    def calc_w(self, p, lineid=0):
        wrest, f, gamma = self.linepars[lineid]
        dl = np.diff(self.obj_wl)
        dl = np.insert(dl, 0, dl[0])
        s, wtot = self.startind, 0.0
        for k in range(int(p[s])):
            logN, z, b = p[1 + 3 * k + s:4 + 3 * k + s]
            t = self._transmission(logN, z, b, (wrest, f, gamma))
            wtot += np.sum((1.0 - t) * dl) / (1.0 + z)
        return float(wtot)

    def calc_N(self, p):
        s, n = self.startind, int(p[self.startind])
        return float(np.log10(np.sum(10.0 ** np.asarray(p[1 + s:1 + s + 3 * n:3])))) if n else -np.inf

    def lnlhood_batch(self, P):
        return np.array([self.lnlhood_worker(p) for p in np.atleast_2d(P)])


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs as concrete inputs live in mcalf_b200.workloads (pure
# numpy input generation, no model code) so that bench.py's GPU arm never imports this module; they
# are re-exported here so oracle, CUDA path and reference all see identical fp64 inputs.
# ---------------------------------------------------------------------------------------------
from mcalf_b200.workloads import config_kwargs, synthetic_spectrum  # noqa: E402,F401


def prior_draws(fitter, B, seed):
    """B parameter vectors: uniform unit cube pushed through ``_scale_cube_pc``."""
    U = np.random.default_rng(seed).random((B, fitter.ndim))
    return np.array([fitter._scale_cube_pc(u) for u in U])
