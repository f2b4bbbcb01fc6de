"""The fitter constructions behind every tag in tests/golden/reference_outputs.npz
(mirrors oracle/make_golden.py so oracle, CUDA path and reference see identical inputs)."""
import os

import numpy as np

from oracle import mcalf_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case(tag):
    """-> (spectrum arrays, ctor kwargs, extra) for a golden tag."""
    spec1, kw1 = orc.config_kwargs(1, GOLDEN)
    if tag in ("cfg1", "cfg1_truth", "edge_strong"):
        return spec1, kw1, {}
    if tag == "cfg2":
        return orc.config_kwargs(2, GOLDEN) + ({},)
    if tag == "cfg2_truth":
        spec2, _ = orc.config_kwargs(2, GOLDEN)
        return spec2, dict(fitrange=[(6180, 6220)], fitlines=["CIV 1548", "CIV 1550"], ncomp=(10, 10),
                           specres=[8.0]), {}
    if tag == "cfg3":
        return orc.config_kwargs(3) + ({},)
    if tag == "cfg4":
        return orc.config_kwargs(4) + ({},)
    wave, flux, err = (a.copy() for a in spec1)
    if tag == "edge_nan":
        flux[100] = np.nan
        err[200] = 0.0
        err[300] = np.nan
        return (wave, flux, err), kw1, {}
    if tag == "edge_ncomp":
        return spec1, dict(kw1, ncomp=(0, 3), contval=[0.8, 1.2]), {}
    if tag == "edge_noconv":
        return spec1, dict(kw1, specres=[0.5, 12.0]), {}
    if tag == "edge_gap":
        return spec1, dict(fitrange=[(6185, 6190), (6200, 6205)], fitlines=["CIV 1548", "CIV 1550"],
                           ncomp=(2, 2), nfill=1, specres=[10.0], contval=[1.0]), {}
    if tag == "edge_asym":
        return spec1, dict(kw1, Asymmlike=True), {"gauss_cdf": (3, 0, 0)}
    raise KeyError(tag)


ALL_TAGS = ["cfg1_truth", "cfg1", "cfg2_truth", "cfg2", "cfg3", "cfg4", "edge_nan", "edge_ncomp",
            "edge_noconv", "edge_gap", "edge_asym", "edge_strong"]
