import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import torch, quick_bench as qb
g=qb.make(4); B=32768
U=torch.rand((B,g.ndim),dtype=torch.float64,device='cuda')
for thr in (256,384,512):
    g.set_option('threads',thr)
    for c in (4,3,2,1):
        g.set_option('ctas_per_sm',c)
        geo=g.geometry()
        ms=qb.timeit(g,U)
        print(thr,geo['ctas_per_sm'],'%.3f ms %.2f M/s'%(ms,B/ms/1e3))
