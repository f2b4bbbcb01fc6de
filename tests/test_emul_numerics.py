"""CPU: the kernels' own arithmetic (mc-alf_b200/csrc/voigt_math.cuh, compiled for the host by
tests/host_emul) against scipy.special.wofz and against the oracle -- the numerics of the fp32 path
can be checked every round without a GPU.  The GPU parity tests (-m gpu) remain the real gate."""
import ctypes

import numpy as np
import pytest
from scipy.special import wofz

from oracle import mcalf_oracle as orc
from tests.cases import case
from tests.host_emul import build as emul_build

dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def emul():
    lib = ctypes.CDLL(emul_build.build())
    lib.emul_epilogue.restype = ctypes.c_double
    return lib


def _call(fn, *arrays):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in arrays]
    out = np.empty_like(arrs[0])
    fn(ctypes.c_long(arrs[0].size), *[a.ctypes.data_as(dp) for a in arrs], out.ctypes.data_as(dp))
    return out


@pytest.mark.parametrize("a0", [1e-5, 1e-4, 1e-3, 1e-2])
def test_fp32_voigt_vs_wofz(emul, a0):
    rng = np.random.default_rng(0)
    u = np.concatenate([rng.uniform(-12, 12, 200000), rng.uniform(-3000, 3000, 50000)]).astype(np.float32).astype(float)
    a = np.full_like(u, np.float32(a0))
    ref = wofz(u + 1j * a).real
    got = _call(emul.emul_voigt_h32, a, u)
    assert (np.abs(got - ref) / ref).max() < 5e-7


def test_fp64_voigt_vs_wofz(emul):
    rng = np.random.default_rng(1)
    for a0 in (1e-6, 1e-4, 1e-2, 0.05, 0.5, 3.0, 20.0):
        u = np.concatenate([rng.uniform(-15, 15, 50000), rng.uniform(-5000, 5000, 20000)])
        a = np.full_like(u, a0)
        ref = wofz(u + 1j * a).real
        assert (np.abs(_call(emul.emul_voigt_h64, a, u) - ref) / ref).max() < 1e-12


def test_depth32(emul):
    x = np.concatenate([np.logspace(-12, 2, 4000), [0.0, 88.0, 500.0]]).astype(np.float32).astype(float)
    got = _call(emul.emul_depth32, x)
    ref = -np.expm1(-x)
    assert got[-3] == 0.0
    assert (np.abs(got - ref) <= 2.5e-7 * ref).all()


def _lines(o, p):
    rows = []
    _, _, nc = o.unpack(p)
    for k in range(nc):
        logN, z, b = p[1 + 3 * k + o.startind:4 + 3 * k + o.startind]
        rows += [(logN, z, b) + lp for lp in o.linepars]
    for k in range(o.nfill):
        logN, z, b = p[3 * k + o.endind:3 * k + 3 + o.endind]
        rows.append((logN, z, b) + o.linefill)
    return np.array(rows, dtype=np.float64).reshape(-1, 6)


@pytest.mark.parametrize("far_eps", [0.0, 3e-9])
@pytest.mark.parametrize("tag", ["cfg1", "cfg2", "cfg3", "cfg4", "edge_gap", "edge_strong"])
def test_kernel_algorithm_on_host_vs_oracle(emul, tag, far_eps, golden):
    """Chunked classification + wing/core forms + depth stencil + two-float residual, run on the host:
    flux within 1e-6 of the continuum and logL within 1e-6 relative of the oracle."""
    spec, kw, extra = case(tag)
    o = orc.OracleFitter(spec, **kw, **extra)
    P = golden[tag + "_P"][:4]
    wave = np.ascontiguousarray(o.obj_wl)
    with np.errstate(all="ignore"):
        w = 1.0 / o.obj_noise ** 2
    seen = set()
    kinds_seen = np.zeros(4, dtype=np.int64)
    for p in P:
        lines = _lines(o, p)
        tau = np.empty(wave.size)
        cls = np.zeros(emul.emul_num_chunks(ctypes.c_long(wave.size), wave.ctypes.data_as(dp)) * max(len(lines), 1), dtype=np.int32)
        kinds = np.zeros(4, dtype=np.int64)
        emul.emul_tau(ctypes.c_long(wave.size), wave.ctypes.data_as(dp), len(lines), lines.ctypes.data_as(dp),
                      ctypes.c_double(0.0), ctypes.c_double(far_eps), tau.ctypes.data_as(dp),
                      cls.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), kinds.ctypes.data_as(ctypes.POINTER(ctypes.c_long)))
        kinds_seen += kinds
        assert set(np.unique(cls)) <= ({1, 2, 3} if far_eps else {1, 2})     # nothing culled at cull_eps = 0
        seen |= set(np.unique(cls))
        res, cont, _ = o.unpack(p)
        model = np.empty(wave.size)
        chi2 = emul.emul_epilogue(ctypes.c_long(wave.size), tau.ctypes.data_as(dp), o.obj.ctypes.data_as(dp),
                                  np.ascontiguousarray(w).ctypes.data_as(dp), ctypes.c_double(res),
                                  ctypes.c_double(o.velstep), ctypes.c_double(cont), model.ctypes.data_as(dp))
        ref_model = o.reconstruct_spec(p)
        assert np.abs(model - ref_model).max() / abs(cont) < 1e-6
        ref_chi2 = o.chi2(p)
        ref_logl = o.lnlhood_worker(p)
        const = ref_logl + 0.5 * ref_chi2
        assert abs((const - 0.5 * chi2) - ref_logl) <= 1e-6 * max(abs(ref_logl), abs(const))
    if far_eps and tag in ("cfg3", "cfg4"):
        assert 3 in seen          # the far-field form is actually exercised
    if tag in ("cfg2", "cfg3", "cfg4"):
        assert kinds_seen[0] > 0 and kinds_seen[1] > 0      # row pairs taken as wing and as core inside mixed chunks
    if tag in ("cfg3", "cfg4"):
        assert kinds_seen[2] > 0                            # ... and pairs straddling table end and core boundary (small b)


@pytest.mark.parametrize("kappa,a0", [(0.01, 1e-4), (0.3, 1.5e-4), (1.0, 3e-4), (8.0, 6e-4), (30.0, 6.5e-4), (200.0, 1e-3)])
def test_weak_line_forms_vs_wofz(emul, kappa, a0):
    """The per-line core boundary (line_cut) and the two forms either side of it: the optical-depth error
    kappa * |H_fp32 - H_wofz| of a line stays below 6e-8 + 3e-7 tau everywhere (flux error F dtau < 1.2e-7)."""
    scut = ctypes.c_double()
    emul.emul_line_cut.restype = ctypes.c_int
    wide = emul.emul_line_cut(ctypes.c_double(kappa), ctypes.c_double(a0), ctypes.byref(scut))
    assert 16.0 <= scut.value <= 36.0
    if not wide:
        assert scut.value == 36.0
        return
    assert scut.value < 35.5 and kappa * np.exp(-scut.value) <= 2.6e-8
    rng = np.random.default_rng(3)
    u = np.concatenate([rng.uniform(-9, 9, 300000), rng.uniform(-400, 400, 50000)]).astype(np.float32).astype(float)
    a = np.full_like(u, np.float32(a0))
    ref = wofz(u + 1j * a).real
    arrs = [np.ascontiguousarray(a), np.ascontiguousarray(u)]
    got = np.empty_like(u)
    emul.emul_voigt_h32_weak(ctypes.c_long(u.size), arrs[0].ctypes.data_as(dp), arrs[1].ctypes.data_as(dp),
                             ctypes.c_double(scut.value), got.ctypes.data_as(dp))
    dtau = kappa * np.abs(got - ref)
    tau = kappa * ref
    if kappa <= 8.0:
        assert (dtau <= 6e-8 + 3e-7 * tau).all(), (dtau - 3e-7 * tau).max()
        assert (np.exp(-tau) * dtau).max() < 1.2e-7
    else:       # strong lines take the two-float core form in the kernel; here only the wing side is theirs
        wing = u * u + a * a >= scut.value
        assert (dtau[wing] <= 6e-8).all()
