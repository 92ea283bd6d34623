"""One launch of the C2 Wagner-Platen fp64 kernel (for ncu): python profiles/prof_c2.py [n_traj] [dt_exp]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stochastic_b200 as pychastic
from stochastic_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296 * 128 * 8
e = int(sys.argv[2]) if len(sys.argv) > 2 else 5
scheme = sys.argv[3] if len(sys.argv) > 3 else "wagner_platen"
if os.environ.get("RECIP"):
    pychastic.config.update("reciprocal_division", bool(int(os.environ["RECIP"])))
prob = workloads.polar_walk(4.0, (2.0, 0.0), 1.0, f64=True)
solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=2.0 ** -e)
if os.environ.get("BLOCK"):
    solver.block_size = int(os.environ["BLOCK"])
if os.environ.get("MINB"):
    solver.min_blocks_per_sm = int(os.environ["MINB"])
if os.environ.get("WS"):
    solver.warp_specialize = bool(int(os.environ["WS"]))
ws = {k: int(os.environ[k]) for k in ("SDE_WS_PGROUPS", "SDE_WS_CONSUMER_REGS", "SDE_WS_PRODUCER_REGS", "SDE_WS_GROUP", "SDE_WS_CTAIL", "SDE_LOGTAB_IN_SMEM", "SDE_WS_SPLIT_NORMAL")
      if os.environ.get(k)}
if ws:
    solver.ws_tuning = ws
if os.environ.get("LOCKSTEP"):
    solver.lockstep = bool(int(os.environ["LOCKSTEP"]))
obs = workloads.polar_error_observables(4.0, 2.0)
for _ in range(2):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    sol = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None, progress_bar=False, observables=obs)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
print(f"n={n} steps={2**e} ms={ms:.3f} traj-steps/s={n * 2**e / ms * 1e3:.4g} regs={solver.last_program.info()}")
