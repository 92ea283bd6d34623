import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from oracle import jax_prng
from stochastic_b200 import runtime
n=1<<20
key=jax_prng.PRNGKey(7)
out=torch.empty(n,dtype=torch.float64,device='cuda')
runtime.check(runtime.lib().sdeb200_normal(int(key[0]),int(key[1]),n,1,0,out.data_ptr(),runtime.current_stream_ptr()))
got=out.cpu().numpy(); ref=jax_prng.normal(key,(n,),np.float64)
ulp=np.abs(got-ref)/np.spacing(np.abs(ref))
print('max ulp',ulp.max(),'at',ref[ulp.argmax()],'mean ulp',ulp.mean(), 'frac>4ulp',(ulp>4).mean())
idx=np.argsort(-ulp)[:10]
for i in idx: print(ref[i], got[i]-ref[i], ulp[i])
rel=np.abs(got-ref)/np.maximum(np.abs(ref),1e-300)
print('max rel',rel.max(), 'max abs', np.abs(got-ref).max())
# f32 failing problems
import problems as P
import stochastic_b200 as pychastic
from oracle import solver as osolver
for name,mk in [('polar',lambda:P.polar_walk(tmax=0.25)),('garch',lambda:P.garch(tmax=0.5))]:
  for scheme in ['milstein','wagner_platen']:
    prob,oprob=mk(); oprob.x0=np.asarray(prob.x0,dtype=np.float64)
    dt=prob.tmax/16
    got=pychastic.sde_solver.SDESolver(scheme=scheme,dt=dt).solve_many(prob,n_trajectories=12,seed=1,progress_bar=False)
    ref=osolver.solve_many(oprob,scheme=scheme,dt=dt,n_trajectories=12,seed=1,dtype=np.float32)
    ref64=osolver.solve_many(oprob,scheme=scheme,dt=dt,n_trajectories=12,seed=1,dtype=np.float64)
    g=np.asarray(got['solution_values'])
    print(name,scheme,'gpu32 vs oracle32',np.abs(g-ref['solution_values']).max(),'oracle32 vs oracle64',np.abs(ref['solution_values']-ref64['solution_values']).max(),'gpu32 vs oracle64',np.abs(g-ref64['solution_values']).max(), 'scale',np.abs(ref['solution_values']).max())
