"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import shim that lets the UNMODIFIED reference package (``/root/reference/mcalf``) be imported in a
container that has numpy + scipy but neither astropy nor linetools.  It injects minimal stand-ins
for the handful of third-party symbols ``mcalf/routines/hires_fitter.py:2-12`` imports; every line
of the reference's own likelihood code (``als_fitter.__init__``, ``voigt_tau``, ``reconstruct_spec``,
``convolve_model``, ``lnlhood_worker``, ``_scale_cube_pc``) then executes as shipped, with the real
``scipy.special.wofz``.

Usable where ``/root/reference`` exists (the build container) or where the pip-installed copy
``baseline/_ref`` travelled along (the GPU box; bench.py's CPU legs).  It is used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/`` and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference tree is absent, e.g. on the GPU
box).  The stand-ins are pinned by the reference's own mock spectra: ``Flux - N(0, 0.02; seed 42)``
of ``testdata/civ_mock_spec*.txt`` equals the model at the truth parameters to 1e-15,
which can only hold if the stand-in ``convolve``/``Gaussian1DKernel``/``sigma_clipped_stats`` and the
CIV atomic constants agree with the real astropy/linetools that produced those files.
"""
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# the reference tree in the build container, else the pip-installed copy that travels to the GPU box
# (`pip install --no-deps --target baseline/_ref /root/reference`: git-ignored, never part of the history)
_CANDIDATES = [os.environ.get("MCALF_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
REFERENCE_ROOT = next((c for c in _CANDIDATES if c and os.path.isdir(os.path.join(c, "mcalf"))), "/root/reference")

# wrest [Angstrom], f, gamma [1/s].  CIV rows are pinned by the golden vectors; the others are
# from memory of Morton (2003) and are UNVERIFIED (harmless: oracle and product share the table).
ATOMIC = {
    "CIV 1548": (1548.204, 0.1899, 2.643e8),
    "CIV 1550": (1550.781, 0.09475, 2.628e8),
    "HI 1215": (1215.67, 0.4164, 6.265e8),
    "HI 1025": (1025.7222, 0.07912, 1.897e8),
    "HI 972": (972.5367, 0.0290, 8.127e7),
    "SiIV 1393": (1393.7602, 0.513, 8.80e8),
    "SiIV 1402": (1402.7729, 0.254, 8.62e8),
}


class _Quantity:
    """number-with-.value, enough for ``x * u.angstrom`` and ``x / u.s`` (hires_fitter.py:104-121)."""

    def __init__(self, value):
        self.value = value

    def __rmul__(self, other):
        return _Quantity(other * self.value)

    def __rtruediv__(self, other):
        return _Quantity(other / self.value)


def _ascii_read(filename):
    with open(filename) as fh:
        names = fh.readline().lstrip("#").split()
    data = np.loadtxt(filename, ndmin=2)
    return {name: data[:, i] for i, name in enumerate(names)}


class _Gaussian1DKernel:
    def __init__(self, stddev, x_size=None):
        half = (int(x_size) - 1) // 2
        x = np.arange(-half, half + 1)
        self.array = np.exp(-0.5 * (x / stddev) ** 2) / (np.sqrt(2.0 * np.pi) * stddev)


def _convolve(array, kernel, boundary="fill", normalize_kernel=True):
    if boundary != "wrap" or not normalize_kernel:
        raise NotImplementedError("only the call made at hires_fitter.py:463 is restated")
    k = kernel.array
    half = len(k) // 2
    a = np.asarray(array, dtype=float)
    padded = np.pad(a, half, mode="wrap")
    out = np.zeros_like(a)
    for j in range(len(k)):
        out += padded[j:j + len(a)] * k[len(k) - 1 - j]
    return out / k.sum()


def _sigma_clipped_stats(data, sigma=3.0, maxiters=5):
    d = np.asarray(data, dtype=float)
    for _ in range(maxiters):
        med = np.median(d)
        keep = np.abs(d - med) <= sigma * d.std()
        if keep.all():
            break
        d = d[keep]
    return d.mean(), np.median(d), d.std()


class _LineList:
    def __init__(self, *args, **kwargs):
        pass

    def __getitem__(self, name):
        if name not in ATOMIC:
            return None
        wrest, f, gamma = ATOMIC[name]
        return {"wrest": _Quantity(wrest), "f": f, "gamma": _Quantity(gamma), "name": name}


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mcalf"))


def install():
    """Inject the stand-ins and return the reference's ``hires_fitter`` module."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    if "astropy" not in sys.modules:
        astropy = mod("astropy")
        io = mod("astropy.io")
        fits = mod("astropy.io.fits")
        asc = mod("astropy.io.ascii")
        units = mod("astropy.units")
        conv = mod("astropy.convolution")
        stats = mod("astropy.stats")
        astropy.io, io.fits, io.ascii, astropy.units = io, fits, asc, units
        units.angstrom, units.s = _Quantity(1.0), _Quantity(1.0)
        asc.read = _ascii_read
        conv.convolve, conv.Gaussian1DKernel = _convolve, _Gaussian1DKernel
        stats.sigma_clipped_stats = _sigma_clipped_stats
    if "linetools" not in sys.modules:
        mod("linetools")
        mod("linetools.lists")
        mod("linetools.lists.linelist").LineList = _LineList
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from mcalf.routines import hires_fitter  # noqa: E402  (the unmodified reference)
    return hires_fitter
