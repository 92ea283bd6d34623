"""European call under a GARCH-diffusion stochastic-volatility model (reference `docs/source/optionpricing.rst:117-167`):
    dS = mu dt + sqrt(v) S dW1,    dv = theta (omega - v) dt + xi v dW2,    S0 = 100, v0 = 0.05, T = 2, strike 120.
Wagner-Platen scheme, the payoff is reduced to (sum, sum of squares) on the device - no path is stored.  With several
GPUs (`torchrun --nproc-per-node N examples/option_pricing.py`) the paths are sharded and one allreduce follows.

    python examples/option_pricing.py [n_paths]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a checkout
import torch.distributed as dist

import stochastic_b200 as pychastic
import stochastic_b200.numpy as jnp
from stochastic_b200.distributed import mean_and_stderr, solve_many_sharded

theta, omega, xi, mu, strike = 0.1, 0.05, 0.1, 0.01, 120.0
n_paths = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 7

if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")

pychastic.config.update("enable_x64", True)
problem = pychastic.sde_problem.SDEProblem(
    a=lambda x: jnp.array([mu, theta * (omega - x[1])]),
    b=lambda x: jnp.array([[jnp.sqrt(x[1]) * x[0], 0], [0, xi * x[1]]]),
    x0=jnp.array([100.0, 0.05]), tmax=2.0)
solver = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=0.01)
solution = solve_many_sharded(solver, problem, n_paths, seed=0, chunk_size=None, store_trajectories=False,
                              observables=lambda x, w, t: jnp.array([jnp.maximum(x[0] - strike, 0.0)]))
mean, stderr = mean_and_stderr(solution["moments"].torch(), solution["count"])
if not dist.is_initialized() or dist.get_rank() == 0:
    print(f"call price (strike {strike:g}) = {float(mean[0]):.4f} +- {float(stderr[0]):.4f}  ({solution['count']} paths)")
if dist.is_initialized():
    dist.destroy_process_group()
