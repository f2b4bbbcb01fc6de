"""GPU: the parity tests that stress index arithmetic (random problems with odd sizes / gaps, long spectra, rows
longer than the CTA, every golden tag, one-component and damped-line cases) re-run against the BOUNDS-CHECKED
build of the library (libmcalf_b200_check.so, -DMCALF_CHECK): every shared-memory and global index the fp32 kernel
forms is validated on the device and a violation fails the call.  compute-sanitizer is closed on the GPU pool;
this is the suite's memory-safety check."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SELFTEST = r"""
import numpy as np, mcalf_b200
from mcalf_b200 import capi
from tests.cases import case
lib = capi.load()
assert lib.mcalf_is_checked_build() == 1
spec, kw, _ = case("cfg1")
g = mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]), specres=kw["specres"], contval=kw["contval"])
P = np.array([[1.0, 13.8, 3.0, 15.0]])
assert np.isfinite(g.lnlhood_batch(P)).all()              # clean run: no violation
g.set_option("check_selftest", 1)
try:
    g.lnlhood_batch(P)
except capi.McalfError as e:
    assert e.code == capi.E_CUDA and "bounds check failed" in str(e) and "0x80000000" in str(e), str(e)
else:
    raise SystemExit("the checked build did not report the deliberate violation")
g.set_option("check_selftest", 0)
assert np.isfinite(g.lnlhood_batch(P)).all()              # the mask is cleared after it was reported
print("selftest ok")
"""


def _env():
    env = dict(os.environ, MCALF_B200_CHECK="1")
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    return env


def test_checked_build_reports_a_deliberate_violation():
    res = subprocess.run([sys.executable, "-c", SELFTEST], cwd=ROOT, env=_env(), capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "selftest ok" in res.stdout, res.stdout + res.stderr


def test_index_heavy_parity_tests_pass_under_the_checked_build():
    sel = ("random_small_problems or long_spectrum or golden_reference_outputs or rows_longer or onecomp or damped "
           "or prior_draws_vs_oracle or device_tensor_path or launch_geometry or contexts_of_different_size or asymmlike")
    res = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_gpu_hardening.py", "-m", "gpu",
                          "-q", "-x", "-p", "no:cacheprovider", "-k", sel], cwd=ROOT, env=_env(), capture_output=True, text=True,
                         timeout=1500)
    tail = "\n".join(res.stdout.splitlines()[-15:])
    assert res.returncode == 0, tail + res.stderr[-2000:]
    assert " passed" in tail and "failed" not in tail, tail
    print(tail.splitlines()[-1])
