"""First passage of a drifted planar Brownian motion through the line y = 2, integrated in polar coordinates
(the study of reference `examples/hit.py` / `examples_from_paper/polar_random_walk_hitting_time.py`).

Two ways to get the hitting place:
  (a) as in the reference: store every step, then post-process each trajectory on the host (`vmap` here is the
      host-side stand-in of `jax.vmap`, see stochastic_b200/hostutil.py);
  (b) on the device: `first_passage=` locates the crossing inside the stepping kernel with the same linear
      interpolation and nothing but (flag, time, state) per trajectory leaves the GPU.
The exact law of the hitting abscissa is known (Brownian motion with drift), so both are compared with it.

    python examples/hitting_time.py [n_trajectories]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a checkout

import stochastic_b200 as pychastic
import stochastic_b200.numpy as jnp

y_drift, y_barrier = 4.0, 2.0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000

problem = pychastic.sde_problem.SDEProblem(
    a=lambda x: jnp.array([1 / (2 * x[0]) + y_drift * jnp.sin(x[1]), y_drift * jnp.cos(x[1]) / x[0]]),
    b=lambda x: jnp.array([[jnp.cos(x[1]), jnp.sin(x[1])], [-jnp.sin(x[1]) / x[0], jnp.cos(x[1]) / x[0]]]),
    x0=jnp.array([2.0, 0.0]), tmax=2.0)
solver = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=2 ** -8)

# (a) the reference's way: all snapshots, host post-processing
solution = solver.solve_many(problem, min(n, 2000), seed=0, chunks_per_randomization=3)
r, phi = solution["solution_values"][..., 0], solution["solution_values"][..., 1]
cart = np.stack([np.asarray(r * jnp.cos(phi)), np.asarray(r * jnp.sin(phi))], axis=-1)


def process_trajectory(trajectory):
    hit_index = (trajectory[:, 1] > y_barrier).argmax()
    before, after = trajectory[hit_index - 1], trajectory[hit_index]
    weight = (y_barrier - before[1]) / (after[1] - before[1])
    return {"hit_x": (after * weight + before * (1 - weight))[0], "hit_index": hit_index + weight}


host = pychastic.vmap(process_trajectory)(cart)

# (b) on the device, for all n trajectories, nothing stored
device = solver.solve_many(problem, n, seed=0, chunk_size=None, chunks_per_randomization=None, store_trajectories=False,
                           first_passage=lambda x, w, t: x[0] * jnp.sin(x[1]) - y_barrier)
flag = np.asarray(device["hit_flag"])
state = np.asarray(device["hit_state"])[flag]
hit_x = state[:, 0] * np.cos(state[:, 1])
hit_t = np.asarray(device["hit_time"])[flag]

# exact: hitting time of level 2 by a BM with drift 4 is inverse Gaussian(mean 1/2, shape 4); x is N(2, T)
print(f"host post-processing  ({len(host['hit_x'])} paths): mean hit x = {np.mean(host['hit_x']):.4f}")
print(f"device first passage ({flag.sum()} of {n} paths hit): mean hit x = {hit_x.mean():.4f} (exact 2), "
      f"mean hit time = {hit_t.mean():.4f} (exact 0.5), var hit x = {hit_x.var():.4f} (exact 0.5)")
