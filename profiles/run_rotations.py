"""Rotational Brownian motion on SO(3) (`step_post_processing`, reference
examples/examples_from_paper/brownian_rotations.py, dt = 0.01, Euler as there; also Milstein / Wagner-Platen), fp64 and fp32,
10^6 trajectories x 200 steps, final state only: python profiles/run_rotations.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stochastic_b200 as pychastic
from stochastic_b200 import workloads

n = 10 ** 6
for f64 in (True, False):
    for scheme in ("euler", "milstein", "wagner_platen"):
        prob, canon = workloads.brownian_rotations(tmax=2.0, f64=f64)
        solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=0.01)
        nn = n if scheme != "wagner_platen" else n // 4
        best = float("inf")
        for rep in range(3):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            sol = solver.solve_many(prob, n_trajectories=nn, seed=0, chunk_size=None, step_post_processing=canon, progress_bar=False)
            ev1.record()
            torch.cuda.synchronize()
            if rep:
                best = min(best, ev0.elapsed_time(ev1))
        rate = nn * solver.last_steps / best * 1e3
        normals = {"euler": 3, "milstein": 2 * 3 + 2 * 10 * 3, "wagner_platen": 3 * 3 + 2 * 10 * 3}[scheme]
        print(f"brownian rotations {'fp64' if f64 else 'fp32'} {scheme}: {rate:.4g} traj-steps/s, {rate * normals:.4g} normals/s, "
              f"{solver.last_program.info()}", flush=True)
