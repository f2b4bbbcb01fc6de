#!/usr/bin/env python
"""bench.py -- batched Voigt-model logL evaluations per second (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

Workload (BASELINE config 4): synthetic 8192-pixel spectrum, CIV doublet,
20 components (40 Voigt lines per sample), floating spectral resolution and continuum, ndim 63;
B unit-cube parameter vectors per GPU (default 262144, the top of the BASELINE sweep) drawn
uniformly from the prior.  One step = one pass of the likelihood hot path over the batch.

  value   whole-job logL/s with the parameter block already resident in HBM (device pointers,
          CUDA events on the launching stream, max over ranks);
  e2e     the same through the public host-buffer API (als_fitter.lnlhood_batch on pinned numpy
          arrays): H2D of the parameters and D2H of logL inside the timed region;
  roofline  FP32 ALU roofline of the fused kernel (this path is compute bound: SURVEY 8d):
          executed FP32 flop/s of the kernel (per-path op counts x the kernel's own path
          counters) over the FFMA peak measured in the same run;
  cpu_baseline  the oracle port (numpy + scipy.special.wofz, the reference's algorithm) on the
          host cores, bounded sample of the same parameter vectors.

Multi-GPU (torchrun, one rank per GPU): the batch shards by sample, no exchange during the
evaluation; every step ends with the logL gather over NCCL.  scaling = weak (B per GPU fixed).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

METRIC = "voigt_logL_evals_per_sec"
UNIT = "logL/s"
WORKLOAD = "cfg4: 8192 px x 20 comps x 2 lines (CIV doublet), free specres+continuum, ndim 63"

# executed FP32 flops per unit of each path of mcalf_fast_kernel (FMA = 2, add/mul/min/max/rint = 1,
# MUFU = 1); derivation in DESIGN.md section 5
FLOP_PAIR_CLASS = 18     # chunk_class per (line, chunk) pair
FLOP_PAIR_FAR = 74       # farfield_accumulate per far pair (8 coefficients x 3 series)
FLOP_NEAR_EVAL = 14      # direct wing form per (line, pixel) of a wing-only pair: u, s (4), Horner with pre-scaled coefficients (8), fused accumulate (2); rcp is MUFU, not counted
FLOP_MIXED_VOTE = 4      # u and s of every pixel of a mixed (line, chunk) pair, computed before the row pair's vote
FLOP_MIXED_WING = 10     # ... plus Horner and accumulate where the row pair takes the wing form
FLOP_CORE_LEAN = 33      # ... or the short core form: coordinate fix-up 5, u^2 and exponent 2, table index 5, H1 cubic 6, a^2 factors 8, combine 5, kappa and accumulate 2 (EX2, min/max not counted)
FLOP_CORE_PRECISE = 80   # ... or the two-float form (two-float coordinate and u^2, polynomial exp)
FLOP_STRADDLE = 11       # extra per pixel of a row pair that needs both forms (wing value + select)
FLOP_CHUNK = 256         # summing the slots' far-field partials: 8 adds on all 32 lanes, per (sample, chunk)
FLOP_PIXEL = 48          # far-field polynomial (15) + depth32 (23) + residual / chi-square (10); stencil: 2 per padded tap
CHUNK = 256
# SURVEY 8d canonical counts (Weideman-32 core, 3-term asymptotic wing)
CANON_EVAL = 5.0
CANON_CORE, CANON_WING = 242.0, 36.0


KIND_TEXT = {"reference": "the unmodified reference (mcalf.routines.hires_fitter.als_fitter.lnlhood_worker, numpy + scipy.special.wofz; "
                          "astropy / linetools stood in by oracle/refshim.py)",
             "port": "oracle port (numpy + scipy.special.wofz)"}


def cfg4_spectrum():
    from mcalf_b200.workloads import config_kwargs    # input generation only
    return config_kwargs(4, GOLDEN)


def make_fitter(device):
    import mcalf_b200
    spec, kw = cfg4_spectrum()
    return mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]),
                                 **{k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()
                                    if k not in ("fitrange", "fitlines", "ncomp")}, device=device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ---------------------------------------------------------------------------------------------
# CPU legs (oracle port; the only place bench.py executes oracle/ code as a workload)
# ---------------------------------------------------------------------------------------------
_CPU_FITTER = None
_CPU_KIND = None


def _cpu_init():
    """One fitter per worker: the UNMODIFIED reference (baseline/_ref or /root/reference, imported through
    oracle/refshim.py, which only stands in for astropy / linetools) when it is there, else the oracle port."""
    global _CPU_FITTER, _CPU_KIND
    spec, kw = cfg4_spectrum()
    try:
        from oracle import refshim
        if not refshim.available():
            raise ImportError("no reference tree")
        hf = refshim.install()
        fd, path = tempfile.mkstemp(suffix=".txt")
        with os.fdopen(fd, "w") as fh:                     # the reference reads an ASCII table with a header line
            fh.write("# Wave Flux Err\n")
            np.savetxt(fh, np.column_stack(spec), fmt="%.17g")
        _CPU_FITTER = hf.als_fitter(path, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]),
                                    nfill=kw.get("nfill", 0), specres=list(kw["specres"]), contval=list(kw["contval"]),
                                    Nrange=list(kw["Nrange"]), brange=list(kw["brange"]))
        os.unlink(path)
        _CPU_KIND = "reference"
    except Exception:                                      # noqa: BLE001 -- any import / construction problem: use the port
        from oracle import mcalf_oracle as orc
        _CPU_FITTER = orc.OracleFitter(spec, **kw)
        _CPU_KIND = "port"


def _cpu_eval(U):
    f = _CPU_FITTER
    return [f.lnlhood_worker(f._scale_cube_pc(u)) for u in U]


def _cpu_kind(_):
    return _CPU_KIND


def cpu_pool():
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init)
    pool.map(_cpu_eval, [np.zeros((0, 63))] * cores)   # every worker builds its fitter outside the timed region
    kinds = set(pool.map(_cpu_kind, range(4 * cores)))
    pool.kind = "reference" if kinds == {"reference"} else "port"
    return pool, cores


def cpu_time(pool, cores, U):
    chunks = [c for c in np.array_split(U, cores) if len(c)]
    t0 = time.perf_counter()
    pool.map(_cpu_eval, chunks)
    return time.perf_counter() - t0


def unit_cube(B, ndim, rank):
    return np.random.default_rng(4000 + rank).random((B, ndim))


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port: numpy + scipy wofz, every host
    core) on a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pool, cores = cpu_pool()
    n = max(cores * 32, 512)      # ~20 s of CPU work per step on one core's clock, spread over the cores
    U = unit_cube(n, 63, 0)
    for _ in range(args.warmup):
        cpu_time(pool, cores, U[:cores])
    t = sum(cpu_time(pool, cores, U) for _ in range(args.steps))
    kind = pool.kind
    pool.close()
    v = n * args.steps / t
    sample = "%d prior-draw parameter vectors of the cfg-4 workload per step, %s" % (n, KIND_TEXT[pool.kind])
    emit({
        "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": n},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": pool.kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def make_cfg_fitter(cfg, device):
    import mcalf_b200
    from mcalf_b200.workloads import config_kwargs
    spec, kw = config_kwargs(cfg, GOLDEN)
    return mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]),
                                 **{k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()
                                    if k not in ("fitrange", "fitlines", "ncomp")}, device=device)


def device_rate(g, U_dev, steps, warmup=2):
    """logL/s with the parameter block resident in HBM: CUDA events on the launching (current torch) stream."""
    import torch
    for _ in range(warmup):
        g.lnlhood_batch(U_dev, unit_cube=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.lnlhood_batch(U_dev, unit_cube=True)
    e1.record()
    torch.cuda.synchronize()
    return U_dev.shape[0] * steps / (e0.elapsed_time(e1) * 1e-3)


def host_rate(g, U_np, out_np, steps, warmup=1):
    """logL/s through the C-ABI with HOST buffers (H2D of the parameters and D2H of logL inside the call)."""
    from mcalf_b200 import capi
    B = U_np.shape[0]

    def call():
        capi.check(g._lib.mcalf_loglike_batch(g._ctx, capi.ptr(U_np), B, g.ndim, capi.F_UNIT_CUBE, None, capi.ptr(out_np), None))

    for _ in range(warmup):
        call()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    return B * steps / (time.perf_counter() - t0)


def far_fraction(g, U_dev):
    g.set_option("collect_stats", 1)
    g.reset_stats()
    g.lnlhood_batch(U_dev, unit_cube=True)
    st = g.stats()
    g.set_option("collect_stats", 0)
    return st["evals_far"] / max(st["evals_total"], 1), st


def run_sweep(g4, local):
    """The rest of the BASELINE sweep, measured in the driver-run line: cfg 4 at every batch size,
    device-resident and end to end (pinned and pageable host buffers), and one line each for cfg 1-3."""
    import torch
    out = {"cfg4": [], "configs": []}
    for B in (4096, 16384, 65536, 262144):
        Uh = torch.from_numpy(unit_cube(B, g4.ndim, 0)).pin_memory()
        Oh = torch.empty(B, dtype=torch.float64).pin_memory()
        Ud = Uh.cuda()
        steps = max(3, min(40, (1 << 20) // B))
        rec = {"batch": B, "device": device_rate(g4, Ud, steps), "e2e_pinned": host_rate(g4, Uh.numpy(), Oh.numpy(), steps)}
        Up = np.array(Uh.numpy())                       # pageable copies: what a CPU sampler hands over
        rec["e2e_pageable"] = host_rate(g4, Up, np.empty(B), steps)
        out["cfg4"].append(rec)
    for cfg, B in ((1, 65536), (2, 65536), (3, 65536)):
        g = make_cfg_fitter(cfg, local)
        Uh = torch.from_numpy(np.random.default_rng(4000 + cfg).random((B, g.ndim))).pin_memory()
        Oh = torch.empty(B, dtype=torch.float64).pin_memory()
        Ud = Uh.cuda()
        ff, st = far_fraction(g, Ud)
        geo = g.geometry()
        out["configs"].append({"cfg": cfg, "batch": B, "npix": geo["npix"], "ndim": g.ndim, "device": device_rate(g, Ud, 6),
                               "e2e_pinned": host_rate(g, Uh.numpy(), Oh.numpy(), 4), "frac_far": ff,
                               "evals_per_logL": st["evals_total"] / B, "cta_threads": geo["threads"], "ctas_per_sm": geo["ctas_per_sm"]})
        g.close()
    out["unit"] = UNIT
    return out


def run_ours(args):
    import hashlib
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus %d must be launched with torchrun (one rank per GPU)" % args.gpus)
        args.gpus = world
    # the CPU-baseline workers are forked before CUDA is initialised in this process (fork after CUDA
    # initialisation is unsafe); they sleep until the GPU measurements are done
    cpu_workers = cpu_pool() if (rank == 0 and world == 1 and not args.no_cpu) else None
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = make_fitter(local)
    B, K, W = args.batch, args.steps, args.warmup
    geo = g.geometry()

    U_host = torch.from_numpy(unit_cube(B, g.ndim, rank)).pin_memory()
    U_dev = U_host.cuda()
    gathered = torch.empty(world * B, dtype=torch.float64, device="cuda") if world > 1 else None

    def step_device():
        logl = g.lnlhood_batch(U_dev, unit_cube=True)
        if world > 1:
            dist.all_gather_into_tensor(gathered, logl)     # the logL gather over NVLink
            return gathered
        return logl

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        out = step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    g.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        out = step_device()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    launches = g.stats()["kernel_launches"]
    checksum = float(out.sum().item())

    # ---- end to end through the host-buffer API ----
    out_host = torch.empty(B, dtype=torch.float64).pin_memory()
    Uh, Oh = U_host.numpy(), out_host.numpy()
    from mcalf_b200 import capi

    def step_host():
        capi.check(g._lib.mcalf_loglike_batch(g._ctx, capi.ptr(Uh), B, g.ndim, capi.F_UNIT_CUBE, None, capi.ptr(Oh), None))
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_host.cuda(non_blocking=True))

    for _ in range(max(1, W // 2)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_host()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    clocks = sampler.stop() if rank == 0 else None
    assert np.array_equal(Oh, out[rank * B:(rank + 1) * B].cpu().numpy()), "host and device paths disagree"
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    e2e_s = float(e2e_s.item())

    # ---- strong scaling (BASELINE config 5): ONE global batch, identical on every rank, sharded by sample
    # through the product's own distributed.ShardedLikelihood; the gathered logL vector must be bit-identical
    # to the same batch evaluated on a single GPU (rank 0 checks) ----
    strong = []
    if not args.no_strong:
        from mcalf_b200.distributed import ShardedLikelihood
        sl = ShardedLikelihood(g, gather=args.gather) if world > 1 else None
        for Bg in (16384, 262144):
            Ug = torch.from_numpy(unit_cube(Bg, g.ndim, 0)).cuda()      # the same seed on every rank
            single = g.lnlhood_batch(Ug, unit_cube=True).clone()       # this GPU alone, the whole batch
            steps = 20 if Bg <= 65536 else max(3, min(K, 10))

            def sstep():
                return sl.lnlhood_batch(Ug, unit_cube=True) if sl is not None else g.lnlhood_batch(Ug, unit_cube=True)

            for _ in range(3):
                full = sstep()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(steps):
                full = sstep()
            s1.record()
            barrier()
            sms = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device="cuda")
            same = torch.tensor([1.0 if torch.equal(full, single) else 0.0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(sms, op=dist.ReduceOp.MAX)
                dist.all_reduce(same, op=dist.ReduceOp.MIN)
            digest = hashlib.sha256(full.cpu().numpy().tobytes()).hexdigest()[:16]
            strong.append({"global_batch": Bg, "batch_per_gpu": -(-Bg // world), "value": Bg * steps / (float(sms.item()) * 1e-3),
                           "ms_per_step": float(sms.item()) / steps, "steps": steps,
                           "global_checksum": digest, "sum": float(full.sum().item()),
                           "bit_identical_to_single_gpu": bool(same.item() == 1.0),
                           "gather": (sl.gather_used if sl is not None else "none")})
            assert same.item() == 1.0, "sharded result differs from the single-GPU result (global batch %d)" % Bg

    if rank == 0:
        # ---- roofline of the fused kernel: measured live, kernel alone on the stream ----
        g.set_option("collect_stats", 1)
        g.reset_stats()
        g.lnlhood_batch(U_dev, unit_cube=True)
        torch.cuda.synchronize()
        st = g.stats()
        g.set_option("collect_stats", 0)
        kms = []
        for _ in range(max(3, min(K, 10))):
            g.lnlhood_batch(U_dev, unit_cube=True)
            torch.cuda.synchronize()
            kms.append(g.stats()["last_kernel_ms"])
        # the kernel's launch duration: at N = 1 the timed region IS K back-to-back launches of it on the launching stream
        # (plus the no-op fp64 fix-up launch that follows each), so its per-step time is the kernel's average duration
        # over the timed region; the isolated launches (synchronised one by one) are reported beside it
        kernel_ms_isolated = float(np.mean(kms))
        kernel_ms = ms / K if world == 1 else kernel_ms_isolated
        P = g.prior_transform_batch(U_dev).cpu().numpy()
        sig = P[:, 0] / 2.354820 / g.velstep
        n = np.where(P[:, 0] > g.velstep, np.ceil(3.0348 * sig), 0)
        npix = g.obj_wl.size
        f_core_canon = 2.0 * math.sqrt(111.0) * np.mean(P[:, 5::3][:, :20]) / g.velstep / npix   # SURVEY 8d estimate
        canon = (st["evals_total"] * (CANON_EVAL + f_core_canon * CANON_CORE + (1 - f_core_canon) * CANON_WING)
                 + npix * (8.0 * B + 2.0 * (2 * n + 1).sum()))
        peak_meas = capi.ffma_peak(local)
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        except (OSError, ValueError):
            sm_max = 1965.0
        peak_nominal = geo["sm_count"] * 128 * 2 * sm_max * 1e6 / 1e12
        # executed FP32 flops per logL: (i) the per-path model (op counts read off the source x the kernel's own
        # path counters of THIS run), (ii) the count derived from the committed ncu opcode histogram of the same
        # command (profiles/ncu_opcodes_r02.json: per-instruction thread counts incl. the packed FFMA2 forms)
        flop_model = model_flops(st, geo, B, npix, n)
        counters, traffic = None, None
        try:
            cj = json.load(open(os.path.join(ROOT, "profiles", "ncu_opcodes_r02.json")))
            if cj.get("samples_per_launch") == B:
                counters = {k: cj[k] for k in ("fp32_flop_per_logL", "warp_instructions_per_logL", "fma_pipe_issue_cycles_per_logL",
                                               "fp32_arith_share_of_warp_instructions", "source") if k in cj}
        except (OSError, ValueError, KeyError):
            pass
        ncu = None
        try:   # dram__bytes_read + dram__bytes_write of this kernel at this batch size, from the committed ncu capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if tr.get("batch") == B:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            ncu = {k: tr[k] for k in ("issue_active_pct", "pipe_fma_pct", "pipe_alu_pct", "pipe_xu_pct", "pipe_lsu_pct",
                                      "warp_instructions_per_logL", "source") if k in tr}
        except (OSError, ValueError, KeyError):
            pass
        flop_per_logl = counters["fp32_flop_per_logL"] if counters else flop_model / B
        achieved = flop_per_logl * B / (kernel_ms * 1e-3) / 1e12
        roofline = {
            "bound": "fp32_alu", "achieved": achieved, "peak": peak_meas, "unit": "TFLOP/s", "frac": achieved / peak_meas,
            "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write; algorithmic: %d)" % (B * (g.ndim * 8 + 8)),
            "peak_source": "FFMA-only microbenchmark in this run (MEASURED_PEAKS.json has no FP32 entry)",
            "peak_nominal": peak_nominal, "frac_of_nominal": achieved / peak_nominal,
            "kernel": "mcalf_fast_kernel", "kernel_ms": kernel_ms, "kernel_ms_isolated": kernel_ms_isolated,
            "flop_per_logL": flop_per_logl,
            "flop_source": ("ncu opcode histogram (committed capture of this command) x this run's kernel time" if counters
                            else "per-path model x this run's path counters"),
            "flop_per_logL_model": flop_model / B, "flop_per_logL_counters": counters["fp32_flop_per_logL"] if counters else None,
            "counters": counters,
            "evals_per_logL": st["evals_total"] / B, "frac_wing": st["evals_wing"] / st["evals_total"],
            "frac_mixed": st["evals_mixed"] / st["evals_total"], "frac_core": st["evals_core"] / st["evals_total"],
            "frac_far": st["evals_far"] / st["evals_total"],
            "canonical_flop_per_logL": canon / B, "canonical_tflops": canon / (kernel_ms * 1e-3) / 1e12,
            "canonical_frac_of_nominal": canon / (kernel_ms * 1e-3) / 1e12 / peak_nominal,
            "hbm_bytes_per_logL": g.ndim * 8 + 8,
            "ncu": ncu,     # from the committed capture under profiles/ (not measured in this run)
        }
        sweep = run_sweep(g, local) if (world == 1 and not args.no_sweep) else None
        # ---- CPU baseline: oracle port on the host cores, bounded sample of the same vectors ----
        cpu = None
        if cpu_workers is not None:
            pool, cores = cpu_workers
            ncpu = max(cores * 32, 512)          # ~20 s of CPU work in total
            tcpu = cpu_time(pool, cores, Uh[:ncpu])
            pool.close()
            cpu = {"value": ncpu / tcpu, "unit": UNIT, "cores": cores, "kind": pool.kind,
                   "sample": "first %d parameter vectors of the timed batch, multiprocessing.Pool(%d), %s" % (ncpu, cores, KIND_TEXT[pool.kind])}
        total = world * B * K
        line = {
            "metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": world * B,
                       "l2_policy": "inputs larger than L2 (%.0f MB parameter block per step)" % (B * g.ndim * 8 / 1e6),
                       "cta_threads": geo["threads"], "ctas_per_sm": geo["ctas_per_sm"], "checksum": checksum,
                       "parallelism": "sample-sharded x%d, logL all-gather" % world},
            "e2e": {"value": total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * g.ndim * 8, "d2h_bytes_per_step": B * 8},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "strong": strong,
            "sweep": sweep,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def model_flops(st, geo, B, npix, n):
    """Executed FP32 flops of one launch from the kernel's own path counters (FMA = 2, add/mul = 1; FFMA2 = 4; what the
    FMA pipe executes: MUFU, min/max, compares and the fp64 set-up are not counted).  Cross-checked against the ncu
    opcode histogram (tools/ncu_opcodes.py): the two agree within a few per cent."""
    taps_padded = 2 * (4 * np.ceil(n / 4)) + 4
    pairs = st["evals_total"] / CHUNK
    lean = st["evals_core"] - st["evals_core_precise"]
    mixed_wing = st["evals_mixed"] - st["evals_core"]
    return (FLOP_PAIR_CLASS * pairs + FLOP_PAIR_FAR * st["evals_far"] / CHUNK + FLOP_NEAR_EVAL * st["evals_wing"]
            + FLOP_MIXED_VOTE * st["evals_mixed"] + FLOP_MIXED_WING * mixed_wing + FLOP_CORE_LEAN * lean
            + FLOP_CORE_PRECISE * st["evals_core_precise"] + FLOP_STRADDLE * st["evals_core_straddle"]
            + FLOP_CHUNK * geo["nchunks"] * B + npix * (FLOP_PIXEL * B + 2.0 * taps_padded.sum()))


def emit(line):
    """The one JSON line goes to the real stdout; everything else this process (or NCCL's C code)
    prints to fd 1 was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)          # e.g. "NCCL version ..." must not precede the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=262144, help="parameter vectors per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the batch-size / config sweep (profiling runs)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling (one global batch) records")
    ap.add_argument("--gather", default="auto", choices=["auto", "nccl", "peer"],
                    help="logL gather of the sharded path: NCCL all-gather, or the kernel's own peer stores")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
