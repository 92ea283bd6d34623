import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import stochastic_b200 as pychastic
from stochastic_b200 import workloads
prob = workloads.polar_walk(4.0,(2.0,0.0),1.0,f64=True)
obs = workloads.polar_error_observables(4.0,2.0)
for n,e,kw in [(5,3,{}),(300,4,dict(chunk_size=3, chunks_per_randomization=2)),(1000,5,dict(chunk_size=None))]:
    a = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=2.0**-e)
    b = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=2.0**-e); b.warp_specialize=True
    ra = a.solve_many(prob, n_trajectories=n, seed=1, observables=obs, **kw)
    rb = b.solve_many(prob, n_trajectories=n, seed=1, observables=obs, **kw)
    torch.cuda.synchronize()
    for k in ("time_values","solution_values","wiener_values","moments"):
        x,y=np.asarray(ra[k]),np.asarray(rb[k])
        print(n,e,k,x.shape, np.abs(x-y).max()/max(1e-300,np.abs(x).max()))
print("regs", b.last_program.info())
