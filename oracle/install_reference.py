"""TEST INFRASTRUCTURE: put a pip-installed copy of the UNMODIFIED reference under ``baseline/_ref`` (git-ignored, not
gpurun-ignored: it travels to the GPU box, where ``/root/reference`` does not exist) so that bench.py's CPU legs can
time the reference's own ``als_fitter.lnlhood_worker`` there.  Nothing of it ever enters the repository's history.

    python oracle/install_reference.py        # only does something where /root/reference exists

The package is pure Python; its third-party imports that are absent from the image (astropy, linetools) are stood in
for by ``oracle/refshim.py`` at import time, hence ``--no-deps``.  The source tree is read-only and setuptools
writes build artefacts next to it, so the install runs from a copy under /tmp."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def install(force=False):
    """-> 'installed' | 'present' | 'no reference tree' | 'failed: ...'"""
    if not os.path.isdir(os.path.join(SRC, "mcalf")):
        return "no reference tree"
    if os.path.isdir(os.path.join(DST, "mcalf")) and not force:
        return "present"
    tmp = tempfile.mkdtemp(prefix="mcalf_ref_")
    try:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(SRC, copy)
        shutil.rmtree(DST, ignore_errors=True)
        res = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                              "--find-links", "/opt/wheelhouse", "--target", DST, copy], capture_output=True, text=True)
        if res.returncode != 0:
            return "failed: " + (res.stderr.strip().splitlines() or ["pip error"])[-1]
        return "installed"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
