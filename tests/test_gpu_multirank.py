"""GPU, two or more devices: the sharded path on real hardware -- one process per GPU over NCCL, the product's
own ``distributed.ShardedLikelihood`` with both gathers (the kernel's peer stores, and NCCL all-gather): the
gathered logL vector must be bit-identical to the same global batch evaluated on a single GPU, for every
world size.  Skipped on a one-GPU box (the driver's GPU tier); run with ``gpurun --gpus 2``."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import mcalf_b200
        from mcalf_b200.distributed import ShardedLikelihood
        from mcalf_b200.workloads import config_kwargs
        golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
        spec, kw = config_kwargs(2, golden)
        g = mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]),
                                  **{k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()
                                     if k not in ("fitrange", "fitlines", "ncomp")}, device=rank)
        report = {"rank": rank}
        for B in (1, 7, 4097, 20000):
            U = torch.from_numpy(np.random.default_rng(77).random((B, g.ndim))).cuda()       # the same block on every rank
            single = g.lnlhood_batch(U, unit_cube=True).clone()
            for mode in ("nccl", "auto"):
                sl = ShardedLikelihood(g, gather=mode)
                for rep in range(3):                      # repeated calls: the double-buffered peer path is exercised
                    out = sl.lnlhood_batch(U, unit_cube=True)
                    torch.cuda.synchronize()
                    ok = bool(torch.equal(out, single))
                    report[(B, mode, rep)] = (ok, sl.gather_used)
        q.put(report)
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_sharded_likelihood_nccl_and_peer_stores_bit_identical():
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    reports = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    used = set()
    for rep in reports:
        for key, val in rep.items():
            if key == "rank":
                continue
            ok, gather = val
            assert ok, (rep["rank"], key, gather)
            used.add((key[1], gather))
    print("world", world, "gathers exercised:", sorted(used))
    assert ("nccl", "nccl") in used


def test_fitter_on_a_device_that_is_not_torchs_current_one():
    """a fitter bound to cuda:1 used while torch's current device is cuda:0 -- the stream handed to the library
    must be cuda:1's, and no call may change the calling thread's current device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    import mcalf_b200
    from mcalf_b200.workloads import config_kwargs
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec, kw = config_kwargs(2, golden)
    args = (spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]))
    kws = {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items() if k not in ("fitrange", "fitlines", "ncomp")}
    torch.cuda.set_device(0)
    g0 = mcalf_b200.als_fitter(*args, **kws, device=0)
    g1 = mcalf_b200.als_fitter(*args, **kws, device=1)
    assert torch.cuda.current_device() == 0
    U = np.random.default_rng(3).random((5000, g0.ndim))
    ref = g0.lnlhood_batch(U, unit_cube=True)
    assert np.array_equal(g1.lnlhood_batch(U, unit_cube=True), ref)                  # host path on device 1
    assert torch.cuda.current_device() == 0
    out = g1.lnlhood_batch(torch.from_numpy(U).to("cuda:1"), unit_cube=True)         # device path: cuda:1's current stream
    torch.cuda.synchronize(1)
    assert out.device.index == 1 and np.array_equal(out.cpu().numpy(), ref)
    assert torch.cuda.current_device() == 0
    with pytest.raises(ValueError):
        g1.lnlhood_batch(torch.from_numpy(U).to("cuda:0"))                           # a tensor of the wrong device is refused
    assert np.array_equal(g1.prior_transform_batch(U[:9]), g0.prior_transform_batch(U[:9]))
    assert torch.cuda.current_device() == 0
