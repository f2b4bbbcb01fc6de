"""TEST INFRASTRUCTURE: g++ build of tests/host_emul/emul.cpp (the kernels' arithmetic on the host)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libmcalf_emul.so")
SRC = os.path.join(HERE, "emul.cpp")
INC = os.path.join(ROOT, "mc-alf_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(INC, f) for f in ("voigt_math.cuh", "voigt_tables.inc")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    # -ffp-contract=off: only the explicit fmaf calls fuse, as in the device code's fma32
    res = subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", INC, SRC, "-o", LIB],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stderr)
    return LIB
