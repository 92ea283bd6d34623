"""Traced bead problems (symbolic Cholesky of the mobility) with the factor formed from one reciprocal square root per pivot
(config "library_rsqrt", default) or from a square root and a division: python profiles/run_library_rsqrt.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import stochastic_b200 as pychastic
from stochastic_b200 import workloads

CASES = [("C3 two RPY spheres, fp64 Euler", lambda: workloads.two_spheres(tmax=5.0, f64=True), "euler", 0.01, 10 ** 6),
         ("four beads (jfm), fp64 Euler", lambda: workloads.four_beads(tmax=5.0, f64=True), "euler", 0.01, 10 ** 6),
         ("C3 two RPY spheres, fp64 Milstein", lambda: workloads.two_spheres(tmax=0.64, f64=True), "milstein", 0.01, 2 * 10 ** 5),
         ("C3 two RPY spheres, fp32 Euler", lambda: workloads.two_spheres(tmax=5.0, f64=False), "euler", 0.01, 10 ** 6)]
for name, make, scheme, dt, n in CASES:
    out = {}
    for flag in (False, True):
        pychastic.config.update("library_rsqrt", flag)
        prob = make()
        solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=dt)
        best = float("inf")
        for rep in range(3):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            sol = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None, progress_bar=False)
            ev1.record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, ev0.elapsed_time(ev1))
        out[flag] = (n * solver.last_steps / best * 1e3, np.asarray(sol["solution_values"])[:1000, -1])
    dev = np.abs(out[True][1] - out[False][1]).max() / np.abs(out[False][1]).max()
    print(f"{name}: sqrt + division {out[False][0]:.4g}, rsqrt {out[True][0]:.4g} traj-steps/s ({out[True][0] / out[False][0]:.3f}x), "
          f"max rel deviation of the first 1000 final states {dev:.2e}", flush=True)
