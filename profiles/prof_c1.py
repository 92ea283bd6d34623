"""Snapshot-heavy run (BASELINE config C1 scaled up): GBM, chunk_size=1, every step stored.
python profiles/prof_c1.py [log2_n_traj] [scheme] ; env SNAP_TILE overrides the staging tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stochastic_b200 as pychastic
from stochastic_b200 import workloads

n = 2 ** (int(sys.argv[1]) if len(sys.argv) > 1 else 21)
schemes = [sys.argv[2]] if len(sys.argv) > 2 else ["euler", "milstein", "wagner_platen"]
prob = workloads.gbm(f64=True)
for scheme in schemes:
    solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=2.0 ** -8)
    if os.environ.get("SNAP_TILE"):
        solver.snap_tile = int(os.environ["SNAP_TILE"])
    best = 1e30
    for _ in range(3):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        sol = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=1, progress_bar=False)
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1))
        del sol
    steps = 512
    gb = n * steps * 24 / 1e9
    print(f"{scheme:14s} n=2^{n.bit_length()-1} ms={best:.2f} traj-steps/s={n*steps/best*1e3:.4g} "
          f"stores={gb:.1f} GB -> {gb/best*1e3:.0f} GB/s tile={solver.snap_tile} regs={solver.last_program.info()['registers']}")
