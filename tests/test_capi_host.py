"""CPU: the C-ABI library loads and exports every symbol include/mcalf_b200.h declares, refuses to
compute without a GPU, and the host-side mirror of als_fitter sets up the same problem as the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import mcalf_oracle as orc
from tests.cases import ALL_TAGS, case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_b", os.path.join(ROOT, "mc-alf_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from mcalf_b200 import capi
    return capi.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from mcalf_b200 import capi
    header = open(os.path.join(ROOT, "include", "mcalf_b200.h")).read()
    declared = set(re.findall(r"\b(mcalf_[a-z0-9_]+)\s*\(", header))
    declared -= {"mcalf_problem", "mcalf_stats"}
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.mcalf_abi_version() == capi.ABI_VERSION


def test_struct_layout_matches_header(lib):
    """sizeof(mcalf_problem_t) / sizeof(mcalf_stats_t) as gcc lays them out == the ctypes mirrors."""
    import subprocess
    import tempfile
    from mcalf_b200 import capi
    src = '#include <stdio.h>\n#include "mcalf_b200.h"\nint main(){printf("%zu %zu\\n", sizeof(mcalf_problem_t), sizeof(mcalf_stats_t));return 0;}\n'
    d = tempfile.mkdtemp()
    open(os.path.join(d, "s.c"), "w").write(src)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
    a, b = map(int, subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split())
    assert (a, b) == (ctypes.sizeof(capi.Problem), ctypes.sizeof(capi.Stats))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful where there is no CUDA device")
def test_no_cpu_fallback(lib):
    """Without a GPU the product path must fail loudly, never compute on the host."""
    import mcalf_b200
    from mcalf_b200 import capi
    spec, kw, _ = case("cfg1")
    with pytest.raises(capi.McalfError) as ei:
        mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]),
                              specres=kw["specres"], contval=kw["contval"])
    assert ei.value.code == capi.E_NODEVICE
    with pytest.raises(capi.McalfError):
        capi.voigt_h(np.zeros(4), np.full(4, 1e-3))
    v = ctypes.c_double()
    assert lib.mcalf_ffma_peak(0, ctypes.byref(v)) == capi.E_NODEVICE


def test_argument_validation(lib):
    from mcalf_b200 import capi
    ctx = ctypes.c_void_p()
    assert lib.mcalf_create(None, 0, ctypes.byref(ctx)) == capi.E_INVALID
    p = capi.Problem()
    p.abi_version = 99
    assert lib.mcalf_create(ctypes.byref(p), 0, ctypes.byref(ctx)) == capi.E_INVALID
    assert b"ABI" in lib.mcalf_last_error()
    assert lib.mcalf_loglike_batch(None, None, 4, 4, 0, None, None, None) == capi.E_INVALID
    assert lib.mcalf_set_option(None, b"x", 0.0) == capi.E_INVALID


@pytest.mark.parametrize("tag", ALL_TAGS)
def test_host_setup_matches_oracle(tag, golden):
    """als_fitter's constructor logic (mask, velstep, bounds, indices) without creating a context."""
    import mcalf_b200
    spec, kw, extra = case(tag)
    o = orc.OracleFitter(spec, **kw, **extra)
    g = mcalf_b200.als_fitter.__new__(mcalf_b200.als_fitter)
    g._init_host(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]),
                 **{k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()
                    if k not in ("fitrange", "fitlines", "ncomp")}, **extra)
    assert g.velstep == o.velstep == float(golden[tag + "_velstep"])
    assert g.ndim == o.ndim and g.startind == o.startind and g.endind == o.endind
    assert np.array_equal(np.column_stack([g._blo, g._bhi]), golden[tag + "_bounds"])
    assert np.array_equal(g.obj_wl, o.obj_wl) and np.array_equal(g.obj, o.obj, equal_nan=True)
    assert g.numlines == o.numlines and g.numfitranges == o.numfitranges
    assert g.linefill["wrest"].value == 250.0 and g.linefill["f"] == g.linepars[0]["f"]
    U = golden.get(tag + "_U")
    if U is not None:
        P = golden[tag + "_P"]
        for u, p in zip(U[:8], P[:8]):
            assert np.array_equal(g._scale_cube_pc(u), p)
            c = u.copy()
            out = g._scale_cube_mn(c, g.ndim, g.ndim)
            assert out is c
            c[g.startind] = int(c[g.startind])
            assert np.array_equal(c, p)
    assert g.lnprior(g._blo) == 0 and g.lnprior(g._bhi + 1) == -np.inf


def test_unknown_line_raises():
    import mcalf_b200
    spec, kw, _ = case("cfg1")
    g = mcalf_b200.als_fitter.__new__(mcalf_b200.als_fitter)
    with pytest.raises(ValueError):
        g._init_host(spec, [[6180, 6220]], ["XX 1234"], [1, 1])


def test_checked_twin_and_new_entry_points(lib):
    """The bounds-checked twin of the library is the same ABI (every symbol, same version) and says what it is; the
    sharded entry point validates its arguments before touching a device."""
    import importlib.util
    from mcalf_b200 import capi
    spec = importlib.util.spec_from_file_location("_b2", os.path.join(ROOT, "mc-alf_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    chk = ctypes.CDLL(mod.build(checked=True))
    for name in capi.SIGNATURES:
        assert getattr(chk, name) is not None
    assert chk.mcalf_abi_version() == capi.ABI_VERSION == lib.mcalf_abi_version()
    assert chk.mcalf_is_checked_build() == 1 and lib.mcalf_is_checked_build() == 0
    arr = (ctypes.c_void_p * 2)(None, None)
    assert lib.mcalf_loglike_batch_peers(None, None, 4, 4, capi.F_ON_DEVICE, None, arr, 0) == capi.E_INVALID      # npeers out of range
    assert lib.mcalf_loglike_batch_peers(None, None, 4, 4, capi.F_ON_DEVICE, None, arr, 9) == capi.E_INVALID
    assert lib.mcalf_loglike_batch_peers(None, None, 4, 4, capi.F_ON_DEVICE, None, arr, 2) == capi.E_INVALID      # null peer buffer
    assert b"peer" in lib.mcalf_last_error()
    assert lib.mcalf_loglike_batch_peers(None, None, 4, 4, 0, None, (ctypes.c_void_p * 1)(8), 1) == capi.E_INVALID  # host pointers refused
    assert b"device pointers" in lib.mcalf_last_error()
