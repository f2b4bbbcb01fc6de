"""Build libmcalf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmcalf_b200.so")
CHECK_LIB = os.path.join(HERE, "libmcalf_b200_check.so")
SOURCES = ["mcalf_kernels.cu", "mcalf_api.cu"]
HEADERS = ["mcalf_device.h", "host_setup.h", "voigt_math.cuh", "voigt_tables.inc", os.path.join("..", "..", "include", "mcalf_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmcalf_b200.so cannot be built")


def stale(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, checked=False):
    """Compile the CUDA kernels and the C-ABI into mc-alf_b200/libmcalf_b200.so (``checked``: the bounds-
    asserting -DMCALF_CHECK variant libmcalf_b200_check.so, test infrastructure); returns its path."""
    lib = CHECK_LIB if checked else LIB
    if not force and not stale(lib):
        return lib
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-DMCALF_CHECK"] if checked else []) + (["-Xptxas", "-v"] if verbose else []) + ["-o", lib] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return lib


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build(force=True, checked=True))
