"""BASELINE.json configs as concrete inputs: spectra and constructor arguments.

Pure numpy input generation -- no model or likelihood code -- shared by bench.py, the tests and the
oracle so that every implementation sees identical fp64 inputs.
"""
import os

import numpy as np

C_KMS = 2.9979245e5    # hires_fitter.py:65


def synthetic_spectrum(window_centres, npix_per_window, velstep_kms=1.0, noise=0.02, seed=42):
    """Log-uniform windows, flux 1 + N(0, noise), as SURVEY §8d cfg 3/4."""
    waves = []
    for wc in window_centres:
        i = np.arange(npix_per_window)
        waves.append(wc * np.exp((i - npix_per_window // 2) * velstep_kms / C_KMS))
    wave = np.concatenate(waves)
    rs = np.random.RandomState(seed)
    flux = 1.0 + rs.normal(0.0, noise, size=wave.size)
    err = np.full(wave.size, noise)
    return wave, flux, err


def config_kwargs(cfg, golden_dir=None):
    """Return (spectrum, ctor-kwargs) for BASELINE config ``cfg`` in {1, 2, 3, 4}."""
    if cfg in (1, 2):
        name = "civ_mock_spec.npz" if cfg == 1 else "civ_mock_spec_multicomp.npz"
        d = np.load(os.path.join(golden_dir, name))
        spectrum = (d["wave"], d["flux"], d["err"])
        if cfg == 1:
            kw = dict(fitrange=[(6180, 6220)], fitlines=["CIV 1548", "CIV 1550"], ncomp=(1, 1),
                      specres=[8.0], contval=[1.0])
        else:
            kw = dict(fitrange=[(6180, 6220)], fitlines=["CIV 1548", "CIV 1550"], ncomp=(8, 11),
                      nfill=2, specres=[8.0, 9.0], contval=[1.0], Nrange=(12.0, 14.5),
                      brange=(10.0, 40.0), zrange=(2.99, 3.01), Nrangefill=(11.5, 16),
                      brangefill=(1, 30))
        return spectrum, kw
    if cfg == 3:
        centres = [4102.9, 4862.7, 5575.0, 6193.0]
        spectrum = synthetic_spectrum(centres, 2048)
        wave = spectrum[0]
        fitrange = []
        for w in range(4):
            seg = wave[w * 2048:(w + 1) * 2048]
            half = 0.5 * (seg[1] - seg[0])
            fitrange.append((seg[0] - half, seg[-1] + half))
        kw = dict(fitrange=fitrange,
                  fitlines=["HI 1215", "HI 1025", "HI 972", "CIV 1548", "CIV 1550", "SiIV 1393", "SiIV 1402"],
                  ncomp=(12, 12), nfill=2, specres=[6.0, 10.0], contval=[0.9, 1.1],
                  Nrange=(12.0, 14.5), brange=(5.0, 40.0), zrange=(2.998, 3.002))
        return spectrum, kw
    if cfg == 4:
        i = np.arange(8192)
        wave = 6180.0 * np.exp(i * 1.0 / C_KMS)
        rs = np.random.RandomState(42)
        flux = 1.0 + rs.normal(0.0, 0.02, size=wave.size)
        err = np.full(wave.size, 0.02)
        kw = dict(fitrange=[(wave[0] - 1.0, wave[-1] + 1.0)], fitlines=["CIV 1548", "CIV 1550"],
                  ncomp=(20, 20), nfill=0, specres=[6.0, 10.0], contval=[0.9, 1.1],
                  Nrange=(12.0, 14.5), brange=(5.0, 40.0))
        return (wave, flux, err), kw
    raise ValueError(cfg)


