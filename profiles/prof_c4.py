"""C4: 40-bead DNA loop (d = m = 120), Euler dt = 1e-7, final state only.
python profiles/prof_c4.py [n_traj] [steps]        (TWIST=1: with the writhe-dependent twist energy; REGTILE=0: shared-memory variant)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stochastic_b200 as pychastic
from stochastic_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
prob = workloads.dna_loop(n_beads=40, tmax=(steps + 0.5) * 1e-7, f64=True, twist=bool(int(os.environ.get("TWIST", "0"))))
solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
if os.environ.get("REGTILE"):
    solver.bead_regtile = bool(int(os.environ["REGTILE"]))
if os.environ.get("BLOCK"):
    solver.block_size = int(os.environ["BLOCK"])
ms = float("inf")
for rep in range(int(os.environ.get("REPS", "4"))):     # the first launch loads the module, the second may still see the clocks ramp up
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    sol = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None, progress_bar=False)
    ev1.record()
    torch.cuda.synchronize()
    if rep:
        ms = min(ms, ev0.elapsed_time(ev1))
rate = n * solver.last_steps / ms * 1e3
print(f"n={n} steps={solver.last_steps} ms={ms:.2f} traj-steps/s={rate:.4g} algorithmic TFLOP/s (0.70 Mflop/step)={rate*0.70e6/1e12:.2f} "
      f"regs={solver.last_program.info()}")
