import sys; sys.path.insert(0, '.')
import numpy as np
from oracle import mile_oracle as o
from mile_b200 import Ensemble, FCNSpec
widths=(256,256,256,256,2); F,N,C=12,700,2
ospec=o.ModelSpec(F,widths,'relu','regr'); rng=np.random.default_rng(0)
X=rng.standard_normal((N,F)).astype(np.float32); y=rng.standard_normal(N).astype(np.float32)
th=(rng.standard_normal((C,ospec.n_params))*(0.5/16)).astype(np.float32)
lp64,g64=o.logpost_batch(ospec,th.astype(np.float64),X.astype(np.float64),y)
for tensor in (2,1,0):
    ens=Ensemble(FCNSpec(F,widths,'relu','regr'),C,tensor=tensor); ens.set_data(X,y)
    lp,g=ens.value_and_grad(th)
    b,k=ospec.offsets(); dims=ospec.dims
    print('tensor',tensor,'lp rel',abs(lp-lp64)/abs(lp64))
    for l in range(5):
        bs=slice(b[l],b[l]+dims[l+1]); ks=slice(k[l],k[l]+dims[l]*dims[l+1])
        r=lambda a,bb: np.linalg.norm(a-bb)/np.linalg.norm(bb)
        print('  layer',l,'bias rel',r(g[0,bs],g64[0,bs]),'kernel rel',r(g[0,ks],g64[0,ks]))
    ens.close()
