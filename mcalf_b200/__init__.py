"""Import name of the ``mc-alf_b200`` package directory (a hyphen cannot be imported).

The sources live in ``mc-alf_b200/`` at the repository root; this stub only extends the package
search path to it, so ``import mcalf_b200`` gives the B200-native MC-ALF likelihood path.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mc-alf_b200"))

from .fitter import als_fitter, ATOMIC  # noqa: E402,F401
from . import capi  # noqa: E402,F401

__all__ = ["als_fitter", "ATOMIC", "capi"]
