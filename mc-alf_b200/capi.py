"""ctypes binding of libmcalf_b200.so (include/mcalf_b200.h) -- the only way the Python layer reaches
the CUDA kernels.  There is no CPU fallback: a missing library or a missing CUDA device raises."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmcalf_b200.so")
CHECK_LIB_PATH = os.path.join(HERE, "libmcalf_b200_check.so")      # -DMCALF_CHECK build (bounds-asserting kernels)

ABI_VERSION = 2
OK, E_INVALID, E_CUDA, E_NODEVICE, E_RESOURCE = 0, -1, -2, -3, -4
F_UNIT_CUBE, F_ON_DEVICE, F_FP64, F_TARGONLY = 0x01, 0x02, 0x04, 0x08
F_ONECOMP, F_ONECOMP_FILL, F_NO_TRUNC, F_FLUX_F64, F_ONELINE = 0x10, 0x20, 0x40, 0x80, 0x100

_dp = ctypes.POINTER(ctypes.c_double)


class Problem(ctypes.Structure):
    """mcalf_problem_t"""
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("npix", ctypes.c_int32),
        ("wave", _dp), ("flux", _dp), ("err", _dp),
        ("velstep", ctypes.c_double),
        ("nlines", ctypes.c_int32),
        ("line_wrest", _dp), ("line_f", _dp), ("line_gamma", _dp),
        ("fill_wrest", ctypes.c_double), ("fill_f", ctypes.c_double), ("fill_gamma", ctypes.c_double),
        ("ncompmax", ctypes.c_int32), ("nfill", ctypes.c_int32),
        ("free_specres", ctypes.c_int32), ("free_cont", ctypes.c_int32),
        ("fixed_specres", ctypes.c_double), ("fixed_cont", ctypes.c_double),
        ("ndim", ctypes.c_int32), ("asymmlike", ctypes.c_int32),
        ("bounds_lo", _dp), ("bounds_hi", _dp),
        ("asym_thresh5", ctypes.c_double), ("asym_thresh4", ctypes.c_double),
        ("max_specres", ctypes.c_double),
    ]


class Stats(ctypes.Structure):
    """mcalf_stats_t"""
    _fields_ = [(n, ctypes.c_uint64) for n in
                ("kernel_launches", "samples", "samples_fp64", "evals_total", "evals_wing", "evals_mixed",
                 "evals_core", "evals_culled", "evals_far", "evals_core_precise", "evals_core_straddle")] + [("last_kernel_ms", ctypes.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/mcalf_b200.h declares: (restype, argtypes)
_vp, _i64, _u32, _int = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint32, ctypes.c_int
SIGNATURES = {
    "mcalf_abi_version": (_int, []),
    "mcalf_last_error": (ctypes.c_char_p, []),
    "mcalf_is_checked_build": (_int, []),
    "mcalf_create": (_int, [ctypes.POINTER(Problem), _int, ctypes.POINTER(_vp)]),
    "mcalf_destroy": (None, [_vp]),
    "mcalf_loglike_batch": (_int, [_vp, _vp, _i64, _i64, _u32, _vp, _vp, _vp]),
    "mcalf_loglike_batch_peers": (_int, [_vp, _vp, _i64, _i64, _u32, _vp, ctypes.POINTER(_vp), _int]),
    "mcalf_model_batch": (_int, [_vp, _vp, _i64, _i64, _u32, _vp, _vp]),
    "mcalf_prior_transform_batch": (_int, [_vp, _vp, _i64, _i64, _u32, _vp, _vp]),
    "mcalf_voigt_h": (_int, [_int, _int, _vp, _vp, _i64, _vp]),
    "mcalf_get_stats": (_int, [_vp, ctypes.POINTER(Stats)]),
    "mcalf_reset_stats": (_int, [_vp]),
    "mcalf_set_option": (_int, [_vp, ctypes.c_char_p, ctypes.c_double]),
    "mcalf_get_option": (_int, [_vp, ctypes.c_char_p, _dp]),
    "mcalf_get_geometry": (_int, [_vp, ctypes.POINTER(_i64)]),
    "mcalf_ffma_peak": (_int, [_int, _dp]),
    "mcalf_host_alloc": (_int, [ctypes.POINTER(_vp), ctypes.c_uint64]),
    "mcalf_host_free": (_int, [_vp]),
}

_lib = None


class McalfError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libmcalf_b200: %s (code %d)" % (message, code))
        self.code = code


def load():
    """Load the shared library and bind every exported symbol; raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = CHECK_LIB_PATH if os.environ.get("MCALF_B200_CHECK") == "1" else LIB_PATH
    path = os.environ.get("MCALF_B200_LIB", path)        # development: an experimental build of the same ABI
    if not os.path.exists(path):
        raise ImportError("%s is missing: build it with `python -m mcalf_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mcalf_abi_version() != ABI_VERSION:
        raise ImportError("libmcalf_b200.so has ABI %d, binding expects %d" % (lib.mcalf_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise McalfError(rc, load().mcalf_last_error().decode("utf-8", "replace"))


def ptr(x):
    """Address of a C-contiguous float64 numpy array, a torch CUDA tensor, an int address or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError("cannot take the address of %r" % type(x))


def voigt_h(u, a, mode=0, device=0):
    """Re w(u + i a) through the kernels' own device code (mode 0: fp32 path, 1: fp64 check path)."""
    lib = load()
    u = np.ascontiguousarray(np.broadcast_to(u, np.broadcast(u, a).shape), dtype=np.float64)
    a = np.ascontiguousarray(np.broadcast_to(a, u.shape), dtype=np.float64)
    out = np.empty_like(u)
    check(lib.mcalf_voigt_h(device, mode, ptr(u), ptr(a), u.size, ptr(out)))
    return out


def ffma_peak(device=0):
    lib = load()
    v = ctypes.c_double()
    check(lib.mcalf_ffma_peak(device, ctypes.byref(v)))
    return v.value
