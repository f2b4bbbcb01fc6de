"""Sample-sharded evaluation over the GPUs of one box.

Every parameter vector is independent, so a batch shards by sample with no exchange during the
evaluation: rank r (one process per GPU, its own ``als_fitter`` context replica) evaluates the
contiguous block ``shard_bounds(B, world)[r]`` and the only collective is the gather of the logL
vector (8 bytes per sample) over NCCL/NVLink -- or gloo on CPU tensors, which is how the host-side
logic is tested without GPUs.  Per-sample results do not depend on the shard a sample lands in
(the kernels' summation orders are fixed by the problem, not by the launch or the batch).
"""
import numpy as np


def shard_bounds(B, world):
    """Contiguous blocks of ceil(B / world) samples; trailing ranks may get fewer (or none)."""
    per = -(-B // world) if B > 0 else 0
    return [(min(r * per, B), min((r + 1) * per, B)) for r in range(world)]


class ShardedLikelihood:
    """``lnlhood_batch`` over all ranks of a ``torch.distributed`` process group.

    ``fitter``   this rank's ``als_fitter`` (bound to this rank's GPU);
    ``evaluate`` optional override ``(rows) -> 1-D tensor`` used instead of ``fitter.lnlhood_batch``
                 (the CPU tests inject a stand-in: the real kernels need a GPU).
    Every rank passes the same full ``[B, ndim]`` block (numpy array or tensor) and receives the full
    ``[B]`` logL vector.
    """

    def __init__(self, fitter=None, group=None, evaluate=None, device=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.fitter = fitter
        self._evaluate = evaluate
        self.device = device
        self.last_shard = None

    def _eval(self, rows, unit_cube):
        if self._evaluate is not None:
            return self._evaluate(rows)
        return self.fitter.lnlhood_batch(rows, unit_cube=unit_cube)

    def lnlhood_batch(self, P, unit_cube=False):
        import torch
        if isinstance(P, np.ndarray):
            P = torch.from_numpy(np.ascontiguousarray(P, dtype=np.float64))
        if P.dim() != 2:
            raise ValueError("expected a [B, ndim] block")
        B = P.shape[0]
        dev = self.device if self.device is not None else (
            torch.device("cuda", self.fitter.device) if self.fitter is not None and self._evaluate is None else P.device)
        lo, hi = shard_bounds(B, self.world)[self.rank]
        self.last_shard = (lo, hi)
        per = -(-B // self.world) if B > 0 else 0
        mine = torch.full((per,), float("nan"), dtype=torch.float64, device=dev)
        if hi > lo:
            rows = P[lo:hi]
            if rows.device != dev:
                rows = rows.to(dev, non_blocking=True)
            out = self._eval(rows, unit_cube)
            if isinstance(out, np.ndarray):
                out = torch.from_numpy(out)
            mine[:hi - lo] = out.to(dev)
        gathered = torch.empty(per * self.world, dtype=torch.float64, device=dev)
        if per:
            self.dist.all_gather_into_tensor(gathered, mine, group=self.group)   # the logL gather
        return gathered[:B]
