import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from scipy.special import wofz
from mcalf_b200 import capi
rng=np.random.default_rng(6)
for a0 in (1e-5,1e-4,1e-3,1e-2):
    u=rng.uniform(-6,6,300000).astype(np.float32).astype(float); a=np.full_like(u,np.float32(a0))
    ref=wofz(u+1j*a).real; got=capi.voigt_h(u,a,mode=2)
    d=np.abs(got-ref); rel=d/ref
    for kappa in (0.1,1.0,8.0):
        sens=kappa*np.exp(-kappa*ref)*d
        print(a0,'kappa',kappa,'max F*dtau %.2e at u=%.2f'%(sens.max(),u[sens.argmax()]),'| max abs %.2e max rel %.2e'%(d.max(),rel.max()))
