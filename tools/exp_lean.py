import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from scipy.special import wofz
from mcalf_b200 import capi
rng=np.random.default_rng(0)
for a0 in (1e-5,1e-4,1e-3,1e-2):
    u=rng.uniform(-6,6,400000).astype(np.float32).astype(float); a=np.full_like(u,np.float32(a0))
    ref=wofz(u+1j*a).real; got=capi.voigt_h(u,a,mode=2)
    rel=np.abs(got-ref)/ref
    print(a0,'lean max rel %.2e'%rel.max(),'at u=%.2f'%u[rel.argmax()],' max abs %.2e'%np.abs(got-ref).max())
