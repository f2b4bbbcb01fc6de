"""Host-side mirror of the reference's likelihood-callback interface, backed by the sm_100a kernels.

``als_fitter`` keeps the constructor signature, the bound methods and the attributes that the
reference's ``cli.py`` and the nested samplers touch (``mcalf/routines/hires_fitter.py:32-34``,
``:202-216``, ``:236-328``, ``:379-449``, ``:697-700``), so a solver that was
handed ``als_fitter.lnlhood_pc`` / ``_dy`` / ``_mn`` and ``_scale_cube_pc`` / ``_mn`` keeps working.
Every model/likelihood evaluation goes through ``libmcalf_b200.so`` (``include/mcalf_b200.h``) via
ctypes; there is no CPU implementation of the hot path in this package.  New, batched entry points
(``lnlhood_batch``, ``chi2_batch``, ``reconstruct_spec_batch``, ``prior_transform_batch``) evaluate
a whole block of live points or proposals per launch.

The constructor (spectrum read, window mask, velocity step, atomic data, prior bounds: the
once-per-run part of the reference, ``hires_fitter.py:65-200``) is plain host Python, as there.
"""
import ctypes
import gc
import os

import numpy as np

from . import capi

C_KMS = 2.9979245e5        # hires_fitter.py:65
FILLER_WREST = 250.0       # hires_fitter.py:121

# wrest [Angstrom], f, gamma [1/s].  linetools' ISM list is used instead when it is installed
# (hires_fitter.py:90).  The CIV rows are pinned by the reference's mock spectra; the others are
# quoted from memory of Morton (2003) and are unverified.
ATOMIC = {
    "CIV 1548": (1548.204, 0.1899, 2.643e8),
    "CIV 1550": (1550.781, 0.09475, 2.628e8),
    "HI 1215": (1215.67, 0.4164, 6.265e8),
    "HI 1025": (1025.7222, 0.07912, 1.897e8),
    "HI 972": (972.5367, 0.0290, 8.127e7),
    "SiIV 1393": (1393.7602, 0.513, 8.80e8),
    "SiIV 1402": (1402.7729, 0.254, 8.62e8),
}
# hires_fitter.py:101-110
_CRII = {"CrII 2066": (0.0512, 4.17e8), "CrII 2062": (0.0759, 4.06e8), "CrII 2056": (0.103, 4.07e8)}


class _Q:
    """A number with ``.value``: what ``linepars[i]['wrest']`` is in the reference (an astropy Quantity)."""
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = float(value)

    def __repr__(self):
        return "<%g>" % self.value


def _lookup_line(name, atomic):
    if atomic and name in atomic:
        return atomic[name]
    try:   # the reference's own source of atomic data, when present
        from linetools.lists.linelist import LineList   # pragma: no cover
        global _LINELIST
        try:
            _LINELIST
        except NameError:
            _LINELIST = LineList("ISM", verbose=False)
        rec = _LINELIST[name]
        if rec is not None:
            return (float(rec["wrest"].value), float(rec["f"]), float(rec["gamma"].value))
    except ImportError:
        pass
    return ATOMIC.get(name)


def read_spectrum(specfile, coldef):
    """ASCII table with a ``# Wave Flux Err`` header line (hires_fitter.py:69-72)."""
    with open(specfile) as fh:
        names = fh.readline().lstrip("#").split()
    data = np.loadtxt(specfile, ndmin=2)
    if len(names) != data.shape[1]:
        raise ValueError("%s: header names %r do not match %d columns" % (specfile, names, data.shape[1]))
    cols = {n: data[:, i] for i, n in enumerate(names)}
    return tuple(np.asarray(cols[c], dtype=float) for c in coldef)


def _sigma_clipped_median(values, sigma=3.0, maxiters=5):
    """astropy.stats.sigma_clipped_stats(...)[1]; only the median is used (hires_fitter.py:84-87)."""
    d = np.asarray(values, dtype=float)
    for _ in range(maxiters):
        keep = np.abs(d - np.median(d)) <= sigma * d.std()
        if keep.all():
            break
        d = d[keep]
    return float(np.median(d))


def _is_device_tensor(x):
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda") and x.is_cuda


class als_fitter:
    """Drop-in for ``mcalf.routines.hires_fitter.als_fitter`` on the likelihood path.

    Extra keyword-only arguments (not in the reference): ``device`` (CUDA ordinal; default
    ``LOCAL_RANK`` or 0), ``precision`` ('fp32' fast kernel | 'fp64' check kernel), ``gauss_cdf``
    (inject the Asymmlike thresholds the reference draws from an unseeded RNG, :179-181),
    ``atomic`` (extra ``{name: (wrest, f, gamma)}``).  ``specfile`` may also be a
    ``(wave, flux, err)`` tuple of arrays.
    """

    def __init__(self, specfile, fitrange, fitlines, ncomp, nfill=0, specres=[7.0], contval=[1.0], Nrange=[11.5, 16],
                 brange=[1, 30], zrange=None, Nrangefill=[11.5, 16], brangefill=[1, 30], wrangefill=None,
                 coldef=['Wave', 'Flux', 'Err'], Gpriors=None, Asymmlike=False, debug=False, *,
                 device=None, precision="fp32", gauss_cdf=None, atomic=None):
        self._init_host(specfile, fitrange, fitlines, ncomp, nfill, specres, contval, Nrange, brange, zrange,
                        Nrangefill, brangefill, wrangefill, coldef, Gpriors, Asymmlike, debug, precision=precision,
                        gauss_cdf=gauss_cdf, atomic=atomic)
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = int(device)
        self._create_context()

    def _init_host(self, specfile, fitrange, fitlines, ncomp, nfill=0, specres=[7.0], contval=[1.0], Nrange=[11.5, 16],
                   brange=[1, 30], zrange=None, Nrangefill=[11.5, 16], brangefill=[1, 30], wrangefill=None,
                   coldef=['Wave', 'Flux', 'Err'], Gpriors=None, Asymmlike=False, debug=False, *,
                   precision="fp32", gauss_cdf=None, atomic=None):
        """The once-per-run host part of the reference constructor (hires_fitter.py:32-200)."""
        self.debug = debug
        self.specfile = specfile
        self.fitrange = fitrange
        self.fitlines = fitlines
        self.Gpriors = Gpriors
        self.Asymmlike = Asymmlike
        if self.Asymmlike:
            print("Running asymmetric likelihood")
        self.specres = specres
        self.contval = contval
        self.ncompmin = ncomp[0]
        self.ncompmax = ncomp[1]
        self.nfill = nfill
        self.freecont = len(contval) > 1
        self.freespecres = len(specres) > 1
        self.clight = C_KMS
        self.ccgs = 2.9979245e10
        if precision not in ("fp32", "fp64"):
            raise ValueError("precision must be 'fp32' or 'fp64'")
        self.precision = precision
        self._ctx = None

        if isinstance(specfile, (str, bytes, os.PathLike)):
            obj_wl, obj, obj_noise = read_spectrum(specfile, coldef)
        else:
            obj_wl, obj, obj_noise = (np.asarray(a, dtype=float) for a in specfile)

        okrange = np.zeros_like(obj_wl, dtype=bool)                      # :75-82
        self.numfitranges = len(self.fitrange)
        for i in range(self.numfitranges):
            okrange[(obj_wl > self.fitrange[i][0]) & (obj_wl < self.fitrange[i][1])] = True
        self.obj = np.ascontiguousarray(obj[okrange])
        self.obj_noise = np.ascontiguousarray(obj_noise[okrange])
        self.obj_wl = np.ascontiguousarray(obj_wl[okrange])
        if self.obj_wl.size < 2:
            raise ValueError("fit ranges select fewer than two pixels")
        velsteps = (self.obj_wl[1:] - self.obj_wl[:-1]) / self.obj_wl[1:] * self.clight
        self.velstep = _sigma_clipped_median(velsteps)                    # :84-87

        self.numlines = len(fitlines)
        linepars = []
        for name in self.fitlines:                                       # :93-113
            rec = _lookup_line(name, atomic)
            if rec is None:
                raise ValueError('ERROR: Line {} not found in database. Aborting.'.format(name))
            wrest, f, gamma = rec
            if name in _CRII:
                f, gamma = _CRII[name]
            linepars.append({"wrest": _Q(wrest), "f": float(f), "gamma": _Q(gamma), "name": name})
        self.linepars = linepars
        self.linefill = dict(linepars[0])                                # :120-121
        self.linefill["wrest"] = _Q(FILLER_WREST)

        self.cont_lims = np.array(contval)
        self.res_lims = np.array(specres)
        self.N_lims = np.array(Nrange)
        self.N_lims_fill = np.array(Nrangefill)
        self.b_lims = np.array(brange)
        self.b_lims_fill = np.array(brangefill)

        w0 = self.linepars[0]["wrest"].value
        self.z_lims = []                                                 # :134-149
        for zz in range(self.ncompmax):
            if zrange is None:
                zmin = ((self.fitrange[0][0] + 0.25) / w0) - 1.
                zmax = ((self.fitrange[0][1] - 0.25) / w0) - 1.
            elif len(zrange) == 2:
                zmin, zmax = zrange[0], zrange[1]
            elif len(zrange) >= 2 * self.ncompmax:
                zmin, zmax = zrange[2 * zz + 0], zrange[2 * zz + 1]
            else:
                raise ValueError('Zrange keyword not understood. Aborting.')
            self.z_lims.append(np.array((zmin, zmax)))
        wf = self.linefill["wrest"].value
        self.z_lims_fill = []                                            # :152-166
        for zz in range(self.nfill):
            if wrangefill is None:
                zmin_fill = ((np.min(self.obj_wl) + 0.25) / wf) - 1.
                zmax_fill = ((np.max(self.obj_wl) - 0.25) / wf) - 1.
            elif len(wrangefill) == 2:
                zmin_fill = (wrangefill[0] / wf) - 1.
                zmax_fill = (wrangefill[1] / wf) - 1.
            elif len(wrangefill) == 2 * self.nfill:
                zmin_fill = (wrangefill[2 * zz + 0] / wf) - 1.
                zmax_fill = (wrangefill[2 * zz + 1] / wf) - 1.
            else:
                raise ValueError('Wrangefill keyword not understood. Aborting.')
            self.z_lims_fill.append(np.array((zmin_fill, zmax_fill)))

        self.startind = int(self.freecont) + int(self.freespecres)       # :169-174
        self.endind = self.startind + 3 * self.ncompmax + 1              # :176

        if gauss_cdf is None:                                            # :179-181 (unseeded in the reference too)
            gauss = np.random.normal(size=len(self.obj))
            gauss_cdf = [(gauss > 3).sum(), (gauss > 4).sum(), (gauss > 5).sum()]
        self.gauss_cdf = [int(v) for v in gauss_cdf]
        self.gracenum = 0.01 * len(self.obj)

        self.bounds = []                                                 # :184-198
        if self.freespecres:
            self.bounds.append(self.res_lims)
        if self.freecont:
            self.bounds.append(self.cont_lims)
        self.bounds.append(ncomp)
        for ii in range(self.ncompmax):
            self.bounds.append(self.N_lims)
            self.bounds.append(self.z_lims[ii])
            self.bounds.append(self.b_lims)
        for ii in range(self.nfill):
            self.bounds.append(self.N_lims_fill)
            self.bounds.append(self.z_lims_fill[ii])
            self.bounds.append(self.b_lims_fill)
        self.ndim = len(self.bounds)                                     # :200
        self._blo = np.array([np.min(b) for b in self.bounds], dtype=np.float64)
        self._bhi = np.array([np.max(b) for b in self.bounds], dtype=np.float64)
        self._ptp = self._bhi - self._blo

    # ------------------------------------------------------------------------------------------
    # context
    # ------------------------------------------------------------------------------------------
    def _create_context(self):
        lib = capi.load()
        as_dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))   # noqa: E731
        self._keep = dict(
            wave=np.ascontiguousarray(self.obj_wl, dtype=np.float64),
            flux=np.ascontiguousarray(self.obj, dtype=np.float64),
            err=np.ascontiguousarray(self.obj_noise, dtype=np.float64),
            lw=np.array([lp["wrest"].value for lp in self.linepars], dtype=np.float64),
            lf=np.array([lp["f"] for lp in self.linepars], dtype=np.float64),
            lg=np.array([lp["gamma"].value for lp in self.linepars], dtype=np.float64),
        )
        k = self._keep
        p = capi.Problem()
        p.abi_version = capi.ABI_VERSION
        p.npix = k["wave"].size
        p.wave, p.flux, p.err = as_dp(k["wave"]), as_dp(k["flux"]), as_dp(k["err"])
        p.velstep = self.velstep
        p.nlines = self.numlines
        p.line_wrest, p.line_f, p.line_gamma = as_dp(k["lw"]), as_dp(k["lf"]), as_dp(k["lg"])
        p.fill_wrest = self.linefill["wrest"].value
        p.fill_f = self.linefill["f"]
        p.fill_gamma = self.linefill["gamma"].value
        p.ncompmax, p.nfill = int(self.ncompmax), int(self.nfill)
        p.free_specres, p.free_cont = int(self.freespecres), int(self.freecont)
        p.fixed_specres = float(max(self.specres))                       # :417
        p.fixed_cont = float(self.contval[0])                            # :425
        p.ndim = self.ndim
        p.asymmlike = int(bool(self.Asymmlike))
        p.bounds_lo, p.bounds_hi = as_dp(self._blo), as_dp(self._bhi)
        p.asym_thresh5 = float(self.gauss_cdf[2] + self.gracenum)        # :300
        p.asym_thresh4 = float(self.gauss_cdf[1] + self.gracenum)        # :302
        p.max_specres = 0.0
        ctx = ctypes.c_void_p()
        capi.check(lib.mcalf_create(ctypes.byref(p), self.device, ctypes.byref(ctx)))
        self._ctx = ctx
        self._lib = lib

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.mcalf_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):                                                 # :697-700
        return self

    def __exit__(self, type, value, trace):
        self.close()
        gc.collect()

    def set_option(self, name, value):
        capi.check(self._lib.mcalf_set_option(self._ctx, name.encode(), float(value)))

    def get_option(self, name):
        v = ctypes.c_double()
        capi.check(self._lib.mcalf_get_option(self._ctx, name.encode(), ctypes.byref(v)))
        return v.value

    def stats(self):
        s = capi.Stats()
        capi.check(self._lib.mcalf_get_stats(self._ctx, ctypes.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        capi.check(self._lib.mcalf_reset_stats(self._ctx))

    def geometry(self):
        g = (ctypes.c_int64 * 8)()
        capi.check(self._lib.mcalf_get_geometry(self._ctx, g))
        names = ("npix", "nchunks", "threads", "ctas_per_sm", "sm_count", "smem_bytes", "halo", "lmax")
        return dict(zip(names, [int(v) for v in g]))

    # ------------------------------------------------------------------------------------------
    # batched entry points (new)
    # ------------------------------------------------------------------------------------------
    def _flags(self, unit_cube=False, fp64=None, targonly=False, no_trunc=False):
        fl = 0
        if unit_cube:
            fl |= capi.F_UNIT_CUBE
        if (self.precision == "fp64") if fp64 is None else fp64:
            fl |= capi.F_FP64
        if targonly:
            fl |= capi.F_TARGONLY
        if no_trunc:
            fl |= capi.F_NO_TRUNC
        return fl

    def _rows(self, P, width):
        """-> (array-or-tensor, B, ld, on_device)"""
        if _is_device_tensor(P):
            import torch
            if P.dtype != torch.float64:
                P = P.to(torch.float64)
            if P.dim() == 1:
                P = P.unsqueeze(0)
            if P.stride(-1) != 1:
                P = P.contiguous()
            if P.device.index != self.device:
                raise ValueError("tensor lives on cuda:%s, the fitter on cuda:%d" % (P.device.index, self.device))
            if P.shape[1] < width:
                raise ValueError("rows have %d entries, need %d" % (P.shape[1], width))
            return P, P.shape[0], P.stride(0) if P.shape[0] > 1 else P.shape[1], True
        A = np.ascontiguousarray(np.atleast_2d(np.asarray(P, dtype=np.float64)))
        if A.shape[1] < width:
            raise ValueError("rows have %d entries, need %d" % (A.shape[1], width))
        return A, A.shape[0], A.shape[1], False

    def _stream(self):
        """torch's current stream ON THE FITTER'S DEVICE (not on torch's current device)."""
        import torch
        return torch.cuda.current_stream(self.device).cuda_stream

    def lnlhood_batch(self, P, unit_cube=False, fp64=None, return_chi2=False, no_trunc=False):
        """logL of every row of ``P`` (physical parameters, or unit-cube draws with ``unit_cube``):
        the batched form of ``lnlhood_worker`` (hires_fitter.py:287-328).  A CUDA tensor is evaluated
        in place on the current torch stream and a CUDA tensor is returned; a numpy array goes
        through the pipelined pinned-staging path and a numpy array comes back."""
        rows, B, ld, on_dev = self._rows(P, self.ndim)
        flags = self._flags(unit_cube, fp64, no_trunc=no_trunc)
        if on_dev:
            import torch
            out = torch.empty(B, dtype=torch.float64, device=rows.device)
            chi = torch.empty(B, dtype=torch.float64, device=rows.device) if return_chi2 else None
            capi.check(self._lib.mcalf_loglike_batch(self._ctx, capi.ptr(rows), B, ld, flags | capi.F_ON_DEVICE,
                                                     self._stream(), capi.ptr(out), capi.ptr(chi)))
        else:
            out = np.empty(B, dtype=np.float64)
            chi = np.empty(B, dtype=np.float64) if return_chi2 else None
            capi.check(self._lib.mcalf_loglike_batch(self._ctx, capi.ptr(rows), B, ld, flags, None, capi.ptr(out),
                                                     capi.ptr(chi)))
        return (out, chi) if return_chi2 else out

    def lnlhood_batch_peers(self, P, peer_ptrs, unit_cube=False, fp64=None, no_trunc=False):
        """The sharded form: logL of every row of the CUDA tensor ``P`` (one shard of a global batch) is stored
        by the kernel itself to every address in ``peer_ptrs`` (+ 8 bytes per row) -- the ranks' gather buffers,
        this GPU's and the others' mapped over NVLink (``mcalf_b200.distributed``).  Enqueues on the current
        torch stream; no collective is involved."""
        rows, B, ld, on_dev = self._rows(P, self.ndim)
        if not on_dev:
            raise ValueError("lnlhood_batch_peers takes a CUDA tensor")
        if not 1 <= len(peer_ptrs) <= 8:
            raise ValueError("1 to 8 peer buffers")
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        flags = self._flags(unit_cube, fp64, no_trunc=no_trunc) | capi.F_ON_DEVICE
        capi.check(self._lib.mcalf_loglike_batch_peers(self._ctx, capi.ptr(rows), B, ld, flags, self._stream(), arr, len(peer_ptrs)))

    def chi2_batch(self, P, unit_cube=False, fp64=None):
        return self.lnlhood_batch(P, unit_cube=unit_cube, fp64=fp64, return_chi2=True)[1]

    def reconstruct_spec_batch(self, P, targonly=False, unit_cube=False, fp64=None, dtype=np.float64):
        """Model flux ``[B, npix]`` for every row of ``P``: batched ``reconstruct_spec`` (:409-449)."""
        rows, B, ld, on_dev = self._rows(P, self.ndim)
        return self._model(rows, B, ld, on_dev, self._flags(unit_cube, fp64, targonly=targonly), dtype)

    def reconstruct_onecomp_batch(self, rows5, fill=False, fp64=None, dtype=np.float64):
        """Rows ``[specres, continuum, N, z, b]`` -> flux of one component (:379-406)."""
        rows, B, ld, on_dev = self._rows(rows5, 5)
        fl = self._flags(False, fp64) | (capi.F_ONECOMP_FILL if fill else capi.F_ONECOMP)
        return self._model(rows, B, ld, on_dev, fl, dtype)

    def _model(self, rows, B, ld, on_dev, flags, dtype):
        npix = self.obj_wl.size
        f64 = np.dtype(dtype) == np.float64 if not on_dev else str(dtype).endswith("float64")
        if f64:
            flags |= capi.F_FLUX_F64
        if on_dev:
            import torch
            out = torch.empty((B, npix), dtype=torch.float64 if f64 else torch.float32, device=rows.device)
            capi.check(self._lib.mcalf_model_batch(self._ctx, capi.ptr(rows), B, ld, flags | capi.F_ON_DEVICE,
                                                   self._stream(), capi.ptr(out)))
            return out
        out = np.empty((B, npix), dtype=np.float64 if f64 else np.float32)
        capi.check(self._lib.mcalf_model_batch(self._ctx, capi.ptr(rows), B, ld, flags, None, capi.ptr(out)))
        return out

    def prior_transform_batch(self, U, no_trunc=False):
        """Unit-cube rows -> physical parameters on the GPU: batched ``_scale_cube_pc`` (:202-209), or
        ``_scale_cube_mn`` (:211-216) with ``no_trunc``."""
        rows, B, ld, on_dev = self._rows(U, self.ndim)
        flags = capi.F_NO_TRUNC if no_trunc else 0
        if on_dev:
            import torch
            out = torch.empty((B, self.ndim), dtype=torch.float64, device=rows.device)
            capi.check(self._lib.mcalf_prior_transform_batch(self._ctx, capi.ptr(rows), B, ld, flags | capi.F_ON_DEVICE,
                                                             self._stream(), capi.ptr(out)))
            return out
        out = np.empty((B, self.ndim), dtype=np.float64)
        capi.check(self._lib.mcalf_prior_transform_batch(self._ctx, capi.ptr(rows), B, ld, flags, None, capi.ptr(out)))
        return out

    # ------------------------------------------------------------------------------------------
    # the reference's scalar interface (each call is a batch of one)
    # ------------------------------------------------------------------------------------------
    def _scale_cube_pc(self, cube):                                      # :202-209
        cube2 = np.copy(cube)
        for ii in range(len(cube)):
            cube2[ii] = cube2[ii] * self._ptp[ii] + self._blo[ii]
            if ii == self.startind:
                cube2[ii] = int(cube2[ii])
        return cube2

    def _scale_cube_mn(self, cube, ndim, nparam):                        # :211-216 (in place, no int())
        for ii in range(ndim):
            cube[ii] = cube[ii] * self._ptp[ii] + self._blo[ii]
        return cube

    def lnprior(self, p):
        """log prior density up to the box normalisation (the quantity hires_fitter.py:218-234 returns):
        -inf outside the bounds; inside, 0 plus a Gaussian term for every parameter whose (value, sigma)
        pair in ``Gpriors`` (flat list, two entries per parameter, the string 'none' = no term) is given."""
        theta = np.asarray([p[i] for i in range(self.ndim)], dtype=np.float64)
        if np.any(theta < self._blo) or np.any(theta > self._bhi) or np.any(np.isnan(theta)):
            return -np.inf
        if self.Gpriors is None:
            return 0
        pairs = [(self.Gpriors[2 * i], self.Gpriors[2 * i + 1]) for i in range(self.ndim)]
        idx = [i for i, (m, sd) in enumerate(pairs) if m != 'none' and sd != 'none']
        if not idx:
            return 0
        mu = np.array([float(pairs[i][0]) for i in idx])
        sd = np.array([float(pairs[i][1]) for i in idx])
        return float(np.sum(-0.5 * (((theta[idx] - mu) / sd) ** 2 + np.log(2.0 * np.pi * sd ** 2))))

    def _row(self, p, n=None):
        n = self.ndim if n is None else n
        return np.array([p[x] for x in range(n)], dtype=np.float64)

    def chi2(self, p):                                                   # :236-248
        row = self._row(p)
        cont = row[1 if self.freespecres else 0] if self.freecont else self.contval[0]
        if cont == 0.:                     # the reference's "model identically zero" branch
            return +np.inf, []
        return float(self.chi2_batch(row[None, :])[0])

    def _scalar_buffers(self):
        b = getattr(self, "_sb", None)
        if b is None:
            row, out = np.empty((1, self.ndim), dtype=np.float64), np.empty(1, dtype=np.float64)
            b = self._sb = (row, out, row.ctypes.data, out.ctypes.data)
        return b

    def lnlhood_worker(self, p):                                         # :287-328
        # batch of one through preallocated buffers: this is the call a CPU sampler makes per point
        row, out, prow, pout = self._scalar_buffers()
        try:
            row[0, :] = p
        except (TypeError, ValueError):
            row[0, :] = [p[x] for x in range(self.ndim)]
        capi.check(self._lib.mcalf_loglike_batch(self._ctx, prow, 1, self.ndim, self._flags(), None, pout, None))
        return float(out[0])

    def lnlhood_pc(self, p):                                             # :250-262
        return self.lnlhood_worker(p), []

    def lnlhood_dy(self, p):                                             # :264-272
        return self.lnlhood_worker(p)

    def lnlhood_mn(self, p, ndim, nparam):                               # :274-285 (p is a ctypes double* under MultiNest)
        row, out, prow, pout = self._scalar_buffers()
        row[0, :] = [p[x] for x in range(self.ndim)]
        capi.check(self._lib.mcalf_loglike_batch(self._ctx, prow, 1, self.ndim, self._flags(), None, pout, None))
        return float(out[0])

    def __call__(self, p):                                               # :509-518 (the reference calls a missing self.lnlhood)
        lp = self.lnprior(p)
        if not np.isfinite(lp):
            return -np.inf
        return lp + self.lnlhood_worker(p)

    def reconstruct_spec(self, p, targonly=False):                       # :409-449
        return self.reconstruct_spec_batch(self._row(p)[None, :], targonly=targonly)[0]

    def reconstruct_onecomp(self, specresolution, continuum, N, z, b):   # :379-392
        return self.reconstruct_onecomp_batch(np.array([[specresolution, continuum, N, z, b]], dtype=np.float64))[0]

    def reconstruct_onecomp_fill(self, specresolution, continuum, N, z, b):   # :394-406
        return self.reconstruct_onecomp_batch(np.array([[specresolution, continuum, N, z, b]], dtype=np.float64),
                                              fill=True)[0]

    # ------------------------------------------------------------------------------------------
    # derived quantities (hires_fitter.py:467-505).  The reference's versions index the parameter
    # vector without the ncomp slot (p[3*comp+startind], :482/:499 -- stale since that slot was
    # introduced) and sum over ncompmax components; here the current layout is used (components at
    # 1+3k+startind) and only the int(p[startind]) active components count.
    # ------------------------------------------------------------------------------------------
    def reconstruct_oneline_batch(self, rows6, fp64=None, dtype=np.float64):
        """Rows ``[specres, continuum, N, z, b, line]`` -> flux of ONE line (index into ``linepars``;
        ``numlines`` = the filler line) of one component: the single-line ``voigt_model`` of :369-377."""
        rows, B, ld, on_dev = self._rows(rows6, 6)
        return self._model(rows, B, ld, on_dev, self._flags(False, fp64) | capi.F_ONELINE, dtype)

    def calc_w_batch(self, P, lineid=0, max_rows=4096):
        """Rest-frame equivalent width of line ``lineid`` summed over the active components, per row of
        ``P`` (batched ``calc_w``, :467-491): sum_k sum_i (1 - T_k(lambda_i)) dlambda_i / (1 + z_k), with the
        unconvolved single-line transmission from the CUDA model kernel (this context, MCALF_F_ONELINE)."""
        P = np.ascontiguousarray(np.atleast_2d(np.asarray(P, dtype=np.float64)))
        dl = np.diff(self.obj_wl)
        dl = np.insert(dl, 0, dl[0])                                       # :486-487
        s, nmax = self.startind, self.ncompmax
        nact = np.clip(np.nan_to_num(P[:, s]).astype(np.int64), 0, nmax)
        comps = P[:, 1 + s:1 + s + 3 * nmax].reshape(-1, nmax, 3)
        owner, slot = np.nonzero(np.arange(nmax)[None, :] < nact[:, None])   # (sample, component) of every active one
        rows = np.empty((owner.size, 6), dtype=np.float64)
        rows[:, 0], rows[:, 1], rows[:, 5] = 0.0, 1.0, float(lineid)         # specres 0: no LSF; continuum 1
        rows[:, 2:5] = comps[owner, slot]
        out = np.zeros(P.shape[0])
        for lo in range(0, len(rows), max_rows):
            flux = self.reconstruct_oneline_batch(rows[lo:lo + max_rows])
            w = ((1.0 - flux) * dl[None, :]).sum(axis=1) / (1.0 + rows[lo:lo + max_rows, 3])
            np.add.at(out, owner[lo:lo + max_rows], w)
        return out

    def calc_w(self, p, lineid=0):
        return float(self.calc_w_batch(np.asarray(p, dtype=np.float64)[None, :], lineid)[0])

    def calc_N(self, p):
        """log10 of the summed column density of the active components (:493-505)."""
        p = np.asarray(p, dtype=np.float64)
        s = self.startind
        n = min(max(int(p[s]), 0), self.ncompmax)
        if n == 0:
            return -np.inf
        return float(np.log10(np.sum(10.0 ** p[1 + s:1 + s + 3 * n:3])))

    def get_jax_likelihood(self):                                        # :521-695
        """A jax-callable ``p[ndim] -> logL`` for jaxns (cli.py:237): the CUDA path behind
        ``jax.pure_callback``; under ``vmap`` the whole block of live points arrives as one batch."""
        from .solvers import jax_likelihood
        return jax_likelihood(self)
