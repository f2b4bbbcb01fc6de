"""TEST INFRASTRUCTURE -- regenerate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py

Writes
  tests/golden/civ_mock_spec.npz, civ_mock_spec_multicomp.npz   the reference's two mock spectra
      (wave, flux, err columns of testdata/*.txt, bit-exact) -- its only golden vectors;
  tests/golden/reference_outputs.npz   parameter vectors + logL / chi2 / model flux computed by the
      reference's own als_fitter (imported through oracle/refshim.py) for BASELINE configs 1-4 and
      for the edge cases the tests cover.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from oracle import mcalf_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TRUTH1 = np.array([1.0, 13.8, 3.0, 15.0])
Z10 = [2.999, 2.9995, 3.0, 3.001, 3.0005, 3.0015, 3.002, 3.0025, 3.0035, 3.0039]
N10 = [13.6, 13.0, 13.8, 13.6, 13.2, 13.4, 13.5, 14.0, 14.2, 13.7]
B10 = [17.5, 8.0, 20.0, 25.0, 15.0, 30.0, 10.0, 25.0, 15.0, 20.0]
TRUTH10 = np.array([10.0] + [v for k in range(10) for v in (N10[k], Z10[k], B10[k])])


def write_spec(tmpdir, name, spectrum):
    path = os.path.join(tmpdir, name)
    np.savetxt(path, np.column_stack(spectrum), header="Wave Flux Err")
    return path


def ref_kwargs(kw):
    """oracle-style kwargs -> reference als_fitter kwargs (lists, as cli.py passes them)."""
    out = {}
    for k, v in kw.items():
        if k == "fitrange":
            out[k] = [list(r) for r in v]
        elif k in ("gauss_cdf",):
            continue
        else:
            out[k] = list(v) if isinstance(v, (tuple, list)) else v
    return out


def main():
    hf = refshim.install()
    os.makedirs(GOLD, exist_ok=True)
    tdir = os.path.join(refshim.REFERENCE_ROOT, "testdata")
    for stem in ("civ_mock_spec", "civ_mock_spec_multicomp"):
        d = np.loadtxt(os.path.join(tdir, stem + ".txt"))
        np.savez_compressed(os.path.join(GOLD, stem + ".npz"), wave=d[:, 0], flux=d[:, 1], err=d[:, 2])

    out = {}
    tmp = tempfile.mkdtemp()

    def run_case(tag, spectrum, kw, P, nflux=2, gauss_cdf=None):
        path = write_spec(tmp, tag + ".txt", spectrum)
        np.random.seed(12345)   # the ctor draws an unseeded normal (hires_fitter.py:179)
        f = hf.als_fitter(path, **ref_kwargs(kw))
        if gauss_cdf is not None:
            f.gauss_cdf = list(gauss_cdf)
        P = np.atleast_2d(np.asarray(P, dtype=float))
        out[tag + "_P"] = P
        with np.errstate(all="ignore"):
            out[tag + "_logL"] = np.array([f.lnlhood_worker(p) for p in P])
            out[tag + "_chi2"] = np.array([f.chi2(p) for p in P])
            out[tag + "_flux"] = np.array([f.reconstruct_spec(p) for p in P[:nflux]])
        out[tag + "_velstep"] = np.array(f.velstep)
        out[tag + "_bounds"] = np.array([[np.min(b), np.max(b)] for b in f.bounds], dtype=float)
        return f

    def draws(f, B, seed):
        U = np.random.default_rng(seed).random((B, f.ndim))
        return U, np.array([f._scale_cube_pc(u) for u in U])

    # --- config 1: truth + prior draws (default wide priors: logN to 16, b down to 1 km/s)
    spec1, kw1 = orc.config_kwargs(1, GOLD)
    run_case("cfg1_truth", spec1, kw1, TRUTH1, nflux=1)
    f = hf.als_fitter(write_spec(tmp, "c1.txt", spec1), **ref_kwargs(kw1))
    U, P = draws(f, 32, 1)
    out["cfg1_U"] = U
    run_case("cfg1", spec1, kw1, P, nflux=4)

    # --- config 2 truth (ncomp=[10,10], specres [8.0], nfill 0) on the multicomp mock
    spec2, kw2 = orc.config_kwargs(2, GOLD)
    kw2t = dict(fitrange=[(6180, 6220)], fitlines=["CIV 1548", "CIV 1550"], ncomp=(10, 10), specres=[8.0])
    run_case("cfg2_truth", spec2, kw2t, TRUTH10, nflux=1)
    f = hf.als_fitter(write_spec(tmp, "c2.txt", spec2), **ref_kwargs(kw2))
    U, P = draws(f, 32, 2)
    out["cfg2_U"] = U
    run_case("cfg2", spec2, kw2, P, nflux=4)

    # --- config 3 / 4 (synthetic), a few draws each (the reference takes 40-60 ms per eval)
    for cfg, B in ((3, 8), (4, 8)):
        spec, kw = orc.config_kwargs(cfg)
        f = hf.als_fitter(write_spec(tmp, "c%d.txt" % cfg, spec), **ref_kwargs(kw))
        U, P = draws(f, B, cfg)
        out["cfg%d_U" % cfg] = U
        run_case("cfg%d" % cfg, spec, kw, P, nflux=2)

    # --- edge cases on the cfg-1 spectrum
    wave, flux, err = (a.copy() for a in spec1)
    # (e1) NaN flux pixel + zero-error pixel + NaN-error pixel are dropped by nansum (:294)
    flux_e, err_e = flux.copy(), err.copy()
    flux_e[100] = np.nan
    err_e[200] = 0.0
    err_e[300] = np.nan
    run_case("edge_nan", (wave, flux_e, err_e), kw1, [TRUTH1, [1.0, 14.5, 3.001, 8.0]], nflux=1)
    # (e2) thisncomp = 0 -> model == continuum; ncomp slot non-integer -> int() truncation (:428)
    kw_e2 = dict(kw1, ncomp=(0, 3), contval=[0.8, 1.2])
    run_case("edge_ncomp", (wave, flux, err), kw_e2,
             [[1.05, 0.0, 13.5, 3.0, 10, 13.5, 3.001, 10, 13.5, 3.002, 10],
              [0.95, 2.7, 13.5, 3.0, 10, 13.9, 3.001, 12, 15.5, 3.002, 10],
              [1.00, 3.0, 13.5, 3.0, 10, 13.9, 3.001, 12, 15.5, 3.002, 3.0]], nflux=3)
    # (e3) specres <= velstep -> no convolution (:445); floating specres straddling velstep
    kw_e3 = dict(kw1, specres=[0.5, 12.0])
    run_case("edge_noconv", (wave, flux, err), kw_e3,
             [[0.9, 1.0, 13.8, 3.0, 15.0], [0.9676, 1.0, 13.8, 3.0, 15.0], [11.7, 1.0, 14.3, 2.9995, 6.0]], nflux=3)
    # (e4) two windows 10 A apart: one concatenated periodic array (:75-82, :463); fillers; line near edges
    kw_e4 = dict(fitrange=[(6185, 6190), (6200, 6205)], fitlines=["CIV 1548", "CIV 1550"], ncomp=(2, 2),
                 nfill=1, specres=[10.0], contval=[1.0])
    run_case("edge_gap", (wave, flux, err), kw_e4,
             [[2, 14.0, 2.99502, 12.0, 13.7, 2.99840, 20.0, 13.0, 23.7605, 9.0],
              [2, 15.8, 2.99820, 2.0, 13.7, 2.99515, 28.0, 15.0, 23.8010, 1.5]], nflux=2)
    # (e5) Asymmlike veto with injected thresholds (:296-303)
    kw_e5 = dict(kw1, Asymmlike=True)
    run_case("edge_asym", (wave, flux, err), kw_e5,
             [TRUTH1, [1.0, 14.5, 3.0005, 25.0], [1.0, 15.5, 3.002, 25.0]], nflux=1, gauss_cdf=(3, 0, 0))
    # (e6) strong / damped-ish and very narrow lines at the default prior corners
    run_case("edge_strong", (wave, flux, err), kw1,
             [[1, 16.0, 3.0, 1.0], [1, 16.0, 3.0, 30.0], [1, 11.5, 3.0, 1.0], [1, 15.0, 2.9925, 3.0],
              [1, 15.9, 3.0172, 1.2]], nflux=5)

    np.savez_compressed(os.path.join(GOLD, "reference_outputs.npz"), **out)
    for k in sorted(out):
        if k.endswith("_logL"):
            print(k, out[k][:3])


if __name__ == "__main__":
    main()
