import os, sys
sys.path.insert(0, os.getcwd())
import torch
import stochastic_b200 as pychastic
from stochastic_b200 import workloads
prob = workloads.gbm(f64=True)
n = 2 ** 21
for scheme in ("euler",):
    for minb, tile in ((None, None), (5, ("tma", 8, 2)), (4, ("tma", 8, 2)), (4, ("tma", 16, 3)), (3, ("tma", 16, 3)), (3, ("tma", 8, 2)), (2, ("tma", 16, 3))):
        solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=2.0 ** -8)
        solver.min_blocks_per_sm = minb
        solver.snap_tile = tile
        best = 1e30
        try:
            for rep in range(4):
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                sol = solver.solve_many(prob, n_trajectories=n, seed=0, chunk_size=1, progress_bar=False)
                ev1.record(); torch.cuda.synchronize()
                if rep: best = min(best, ev0.elapsed_time(ev1))
                del sol
            print(scheme, "min_blocks", minb, "tile", tile, "%.3e traj-steps/s %.0f GB/s" % (n * 512 / best * 1e3, n * 512 * 24 / best / 1e6), "regs", solver.last_program.info()["registers"], flush=True)
        except Exception as e:
            print(scheme, minb, tile, "failed:", str(e)[:100])
