import time, numpy as np, torch, sys
sys.path.insert(0,'.')
from mile_b200 import ShardedEnsemble, Ensemble
from mile_b200 import synthetic as syn
key='covertype_full'; C=12
spec=syn.workload_spec(key); X,y,Xt,yt=syn.synthetic_data(key, seed=1234); d=spec.n_params
for cls in (ShardedEnsemble, Ensemble):
    ens = cls(spec, C, device=0, rank=0, world=1) if cls is ShardedEnsemble else cls(spec, C, device=0)
    ens.set_data(X,y); th0=syn.synthetic_theta0(d,C,seed0=1000,scale=0.3); ens.init(th0, seed=17)
    eps=np.full(C,0.02,np.float32); L=np.full(C,np.sqrt(d),np.float32)
    Xp=torch.from_numpy(X).pin_memory().numpy(); st=ens.get_state()
    ens.sample(20,eps,L,n_thinning=10,seed=1)
    for it in range(3):
        t=[time.perf_counter()]
        ens.set_data(Xp,y); t.append(time.perf_counter())
        ens.set_state(*st); t.append(time.perf_counter())
        ens.sample(20,eps,L,n_thinning=10,seed=2+it); t.append(time.perf_counter())
        st=ens.get_state(); t.append(time.perf_counter())
        print(cls.__name__, ' '.join(f'{(b-a)*1e3:8.2f}' for a,b in zip(t,t[1:])), 'ms (set_data set_state sample get_state)')
    ens.close()
