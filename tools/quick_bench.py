"""Scratch timing of the fp32 kernel on one GPU (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mcalf_b200
from mcalf_b200.workloads import config_kwargs

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

def make(cfg):
    spec, kw = config_kwargs(cfg, GOLD)
    return mcalf_b200.als_fitter(spec, [list(r) for r in kw["fitrange"]], kw["fitlines"], list(kw["ncomp"]),
        **{k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items() if k not in ("fitrange", "fitlines", "ncomp")})

def timeit(g, U, reps=3):
    torch.cuda.synchronize()
    g.lnlhood_batch(U, unit_cube=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.lnlhood_batch(U, unit_cube=True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

if __name__ == "__main__":
    cfgs = [int(a) for a in sys.argv[1:]] or [4]
    print("ffma peak TFLOP/s", mcalf_b200.capi.ffma_peak())
    for cfg in cfgs:
        g = make(cfg)
        B = {1: 262144, 2: 65536, 3: 16384, 4: 32768}[cfg]
        U = torch.rand((B, g.ndim), dtype=torch.float64, device="cuda")
        print("cfg", cfg, g.geometry())
        for threads in (0, 128, 256, 512, 1024):
            try:
                g.set_option("threads", threads)
            except Exception as e:
                print("threads", threads, e); continue
            geo = g.geometry()
            ms = timeit(g, U)
            print("cfg %d threads %4d ctas/sm %d smem %6d : %8.3f ms  %10.0f logL/s" % (cfg, geo["threads"], geo["ctas_per_sm"], geo["smem_bytes"], ms, B / ms * 1e3))
        g.set_option("threads", 0)
        g.set_option("collect_stats", 1); g.reset_stats()
        g.lnlhood_batch(U, unit_cube=True); torch.cuda.synchronize()
        print(g.stats())
