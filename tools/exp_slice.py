import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, numpy as np, quick_bench as qb
from mcalf_b200 import capi
g = qb.make(4); B = 262144
Uh = torch.rand((B, g.ndim), dtype=torch.float64).pin_memory(); Oh = torch.empty(B, dtype=torch.float64).pin_memory()
U, O = Uh.numpy(), Oh.numpy()
def run():
    capi.check(g._lib.mcalf_loglike_batch(g._ctx, capi.ptr(U), B, g.ndim, capi.F_UNIT_CUBE, None, capi.ptr(O), None))
for sl in (8192, 16384, 32768, 65536, 131072, 262144):
    g.set_option('slice', sl)
    run(); run()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): run()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print('slice', sl, '%.2f ms  %.2f M/s' % (dt * 1e3, B / dt / 1e6))
Ud = Uh.cuda()
print('device', qb.timeit(g, Ud), 'ms')
