"""One process, several GPUs: a batch is sharded by sample over per-device ``als_fitter`` contexts that
run concurrently on host threads (the ctypes calls release the GIL).  This is for a single-process
sampler that wants the whole box (dynesty + ``BatchPool``, a notebook); multi-process runs
(torchrun / MPI, one rank per GPU) use ``mcalf_b200.distributed`` instead.  No collective is involved:
each context returns its shard's logL to host memory and the shards are concatenated.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .distributed import shard_bounds
from .fitter import als_fitter


class MultiDeviceFitter:
    """``als_fitter`` replicas on ``devices`` (default: every visible GPU) behind the batched interface.
    Scalar callbacks and attributes are served by the first replica."""

    def __init__(self, *args, devices=None, **kwargs):
        if devices is None:
            import torch
            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise RuntimeError("no CUDA device")
        self.devices = list(devices)
        self.fitters = [als_fitter(*args, device=d, **kwargs) for d in self.devices]
        self._pool = ThreadPoolExecutor(max_workers=len(self.fitters))

    def __getattr__(self, name):            # ndim, bounds, lnlhood_pc, _scale_cube_pc, reconstruct_spec, ...
        return getattr(self.fitters[0], name)

    def _sharded(self, method, P, **kw):
        P = np.ascontiguousarray(np.atleast_2d(np.asarray(P, dtype=np.float64)))
        bounds = [(f, lo, hi) for f, (lo, hi) in zip(self.fitters, shard_bounds(P.shape[0], len(self.fitters))) if hi > lo]
        if not bounds:
            return getattr(self.fitters[0], method)(P, **kw)
        parts = list(self._pool.map(lambda b: getattr(b[0], method)(P[b[1]:b[2]], **kw), bounds))
        return np.concatenate(parts, axis=0)

    def lnlhood_batch(self, P, unit_cube=False, fp64=None):
        return self._sharded("lnlhood_batch", P, unit_cube=unit_cube, fp64=fp64)

    def chi2_batch(self, P, unit_cube=False, fp64=None):
        return self._sharded("chi2_batch", P, unit_cube=unit_cube, fp64=fp64)

    def reconstruct_spec_batch(self, P, targonly=False, unit_cube=False, fp64=None, dtype=np.float64):
        return self._sharded("reconstruct_spec_batch", P, targonly=targonly, unit_cube=unit_cube, fp64=fp64, dtype=dtype)

    def prior_transform_batch(self, U, no_trunc=False):
        return self._sharded("prior_transform_batch", U, no_trunc=no_trunc)

    def close(self):
        for f in self.fitters:
            f.close()
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
