import os, sys
sys.path.insert(0, os.getcwd())
import torch
import stochastic_b200 as pychastic
from stochastic_b200 import workloads
def run(name, prob, scheme, dt, n, **kw):
    for recip in (False, True):
        pychastic.config.update("reciprocal_division", recip)
        solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=dt)
        best = 1e30
        for rep in range(4):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            sol = solver.solve_many(prob, n_trajectories=n, seed=0, progress_bar=False, chunk_size=None, **kw)
            ev1.record(); torch.cuda.synchronize()
            if rep: best = min(best, ev0.elapsed_time(ev1))
        print(name, scheme, "recip", recip, "%.3e traj-steps/s" % (n * solver.last_steps / best * 1e3), "regs", solver.last_program.info()["registers"], flush=True)
run("two_spheres", workloads.two_spheres(tmax=5.0, f64=True), "euler", 0.01, 10**6)
run("four_beads", workloads.four_beads(tmax=20.0, f64=True), "euler", 0.05, 10**6)
run("polar", workloads.polar_walk(4.0, (2.0, 0.0), 1.0, f64=True), "milstein", 2**-7, 10**7)
run("polar", workloads.polar_walk(4.0, (2.0, 0.0), 1.0, f64=True), "euler", 2**-7, 10**7)
