"""Five device-resident launches of the fp32 kernel at one BASELINE config (for an ncu capture of that config):
    ncu --set full -k regex:mcalf_fast_kernel -s 2 -c 1 -o prof_cfgN python tools/profile_cfg.py N [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, quick_bench as qb
cfg = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else {1: 262144, 2: 131072, 3: 65536, 4: 262144}[cfg]
g = qb.make(cfg)
U = torch.from_numpy(__import__("numpy").random.default_rng(4000 + cfg).random((B, g.ndim))).cuda()
ms = qb.timeit(g, U, reps=4)
print("cfg %d B %d: %.3f ms per launch, %.2f M logL/s" % (cfg, B, ms, B / ms / 1e3))
