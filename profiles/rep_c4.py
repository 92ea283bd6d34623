import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stochastic_b200 as pychastic
from stochastic_b200 import workloads
prob = workloads.dna_loop(n_beads=40, tmax=100.5e-7, f64=True)
solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
for k in range(10):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    sol = solver.solve_many(prob, n_trajectories=8880, seed=k, chunk_size=None, progress_bar=False)
    ev1.record()
    torch.cuda.synchronize()
    print("launch", k, "event ms", round(ev0.elapsed_time(ev1), 2), flush=True)
