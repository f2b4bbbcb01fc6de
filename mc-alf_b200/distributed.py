"""Sample-sharded evaluation over the GPUs of one box.

Every parameter vector is independent, so a batch shards by sample with no exchange during the
evaluation: rank r (one process per GPU, its own ``als_fitter`` context replica -- the analogue of the
reference's MPI ranks each owning an ``als_fitter``, ``cli.py:37-41,156-158``) evaluates the contiguous
block ``shard_bounds(B, world)[r]``; the only exchange is the gather of the logL vector (8 bytes per
sample).  Two gathers:

* ``"peer"``  the kernel's own tail does it: every rank's gather buffer lives in symmetric memory
  (``torch.distributed._symmetric_memory``: each rank's buffer mapped into every other rank's address
  space over NVLink), and ``mcalf_loglike_batch_peers`` makes the kernel store each sample's logL
  straight into all ranks' buffers at the shard's offset.  No collective follows, only a
  symmetric-memory barrier on the stream.
* ``"nccl"``  ``all_gather_into_tensor`` over NCCL (or gloo on CPU tensors, which is how the host-side
  logic is tested without GPUs).

``gather="auto"`` takes the peer stores when symmetric memory can be set up on every rank and falls
back to NCCL otherwise; ``gather_used`` says which ran.  Per-sample results do not depend on the shard
a sample lands in (the kernels' summation orders are fixed by the problem, not by the launch or the
batch), so the gathered vector is bit-identical for every world size.
"""
import numpy as np


def shard_bounds(B, world):
    """Contiguous blocks of ceil(B / world) samples; trailing ranks may get fewer (or none)."""
    per = -(-B // world) if B > 0 else 0
    return [(min(r * per, B), min((r + 1) * per, B)) for r in range(world)]


class _PeerBuffers:
    """Two gather buffers per rank in symmetric memory (double-buffered: a rank may still be reading call
    k's result while a faster peer's call k+1 already stores into the other half)."""

    def __init__(self, dist, group, device, capacity):
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        self.capacity = int(capacity)
        self.buf = symm_mem.empty(2 * self.capacity, dtype=torch.float64, device=device)
        grp = group if group is not None else dist.group.WORLD
        try:
            self.hdl = symm_mem.rendezvous(self.buf, grp)
        except TypeError:
            self.hdl = symm_mem.rendezvous(self.buf, grp.group_name)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.turn = 0

    def next_half(self):
        half = self.turn
        self.turn ^= 1
        return half


class ShardedLikelihood:
    """``lnlhood_batch`` over all ranks of a ``torch.distributed`` process group.

    ``fitter``   this rank's ``als_fitter`` (bound to this rank's GPU);
    ``evaluate`` optional override ``(rows) -> 1-D tensor`` used instead of ``fitter.lnlhood_batch``
                 (the CPU tests inject a stand-in: the real kernels need a GPU);
    ``gather``   ``"auto"`` | ``"peer"`` | ``"nccl"`` (see the module docstring).
    Every rank passes the same full ``[B, ndim]`` block (numpy array or tensor) and receives the full
    ``[B]`` logL vector.
    """

    def __init__(self, fitter=None, group=None, evaluate=None, device=None, gather="auto"):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if gather not in ("auto", "peer", "nccl"):
            raise ValueError("gather must be 'auto', 'peer' or 'nccl'")
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.fitter = fitter
        self._evaluate = evaluate
        self.device = device
        self.last_shard = None
        self._want = gather
        self._peer = None
        self._peer_failed = None
        self.gather_used = "nccl"

    # ---- the peer-store gather ----
    def _peer_buffers(self, per, dev):
        """Symmetric gather buffers holding ``per * world`` doubles, (re)built collectively when the batch grows.
        Returns None (on every rank alike) when symmetric memory cannot be set up."""
        import torch
        need = per * self.world
        if self._peer is not None and self._peer.capacity >= need:
            return self._peer
        if self._peer_failed:
            return None
        ok, err = 1, None
        try:
            peer = _PeerBuffers(self.dist, self.group, dev, max(need, 1 << 15))
        except Exception as e:            # noqa: BLE001 -- any failure means "use NCCL", but every rank must agree
            ok, err, peer = 0, e, None
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            self._peer_failed = err or RuntimeError("symmetric memory unavailable on another rank")
            if self._want == "peer":
                raise RuntimeError("gather='peer' requested but symmetric memory could not be set up: %r" % (self._peer_failed,))
            return None
        self._peer = peer
        return peer

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
        on_gpu = self.fitter is not None and self._evaluate is None
        dev = self.device if self.device is not None else (torch.device("cuda", self.fitter.device) if on_gpu else P.device)
        lo, hi = shard_bounds(B, self.world)[self.rank]
        self.last_shard = (lo, hi)
        per = -(-B // self.world) if B > 0 else 0
        rows = P[lo:hi]
        if hi > lo and rows.device != dev:
            rows = rows.to(dev, non_blocking=True)

        peer = self._peer_buffers(per, dev) if (on_gpu and self._want != "nccl" and per > 0) else None
        if peer is not None:
            # the kernel stores this shard's logL into EVERY rank's gather buffer (its own included) at offset lo
            self.gather_used = "peer"
            half = peer.next_half()
            base = half * peer.capacity
            if hi > lo:
                self.fitter.lnlhood_batch_peers(rows, [p + 8 * (base + lo) for p in peer.ptrs], unit_cube=unit_cube)
            peer.hdl.barrier(channel=half)          # on the current stream: every rank's stores have landed
            # a private copy: the symmetric buffer half is written again two calls from now (the copy is stream-ordered
            # before this rank arrives at the next barrier, which is what lets peers reuse the half safely)
            return peer.buf[base:base + B].clone()

        self.gather_used = "nccl" if on_gpu else "collective"
        mine = torch.full((per,), float("nan"), dtype=torch.float64, device=dev)
        if hi > lo:
            out = self._eval(rows, unit_cube)
            if isinstance(out, np.ndarray):
                out = torch.from_numpy(out)
            mine[:hi - lo] = out.to(dev)
        gathered = torch.empty(per * self.world, dtype=torch.float64, device=dev)
        if per:
            self.dist.all_gather_into_tensor(gathered, mine, group=self.group)   # the logL gather
        return gathered[:B]
