"""Throughput of the BASELINE configs on one GPU (table for BASELINE.md section 3 / profiles/README.md).
python profiles/run_configs.py [scale]   (scale < 1 shrinks trajectory counts for a quick run)"""
import json, os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import stochastic_b200 as pychastic
from stochastic_b200 import workloads, runtime

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
tf, ms = ctypes.c_double(), ctypes.c_double()
runtime.check(runtime.lib().sdeb200_fp64_peak(8192, tf, ms))
peak = tf.value
rows = []


def run(name, prob, scheme, dt, n, flops, reps=3, **kw):
    solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=dt)
    best = 1e30
    for _ in range(reps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        sol = solver.solve_many(prob, n_trajectories=n, seed=0, progress_bar=False, **kw)
        ev1.record()
        torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1))
    steps = solver.last_steps
    rate = n * steps / best * 1e3
    row = {"config": name, "scheme": scheme, "n_traj": n, "steps": steps, "ms": round(best, 3),
           "traj_steps_per_s": rate, "algorithmic_flops_per_step": flops,
           "algorithmic_tflops": rate * flops / 1e12, "frac_fp64_peak": rate * flops / 1e12 / peak,
           "registers": solver.last_program.info()["registers"]}
    if kw.get("chunk_size", 1) == 1:
        d, m = prob.dimension, prob.noise_terms
        row["store_GBps"] = rate * (1 + d + m) * 8 / 1e9
    if "moments" in sol:
        mom = np.asarray(sol["moments"])
        row["moment_means"] = (mom[:, 0] / sol["count"]).tolist()
    rows.append(row)
    print(json.dumps(row), flush=True)
    return sol


n1 = int(10 ** 4)
for scheme, fl in (("euler", 62), ("milstein", 69), ("wagner_platen", 145)):
    run("C1 GBM 1e4 traj, every step stored", workloads.gbm(f64=True), scheme, 2 ** -8, n1, fl, chunk_size=1)
    run("C1 GBM 2^21 traj, every step stored", workloads.gbm(f64=True), scheme, 2 ** -8, int(2 ** 21 * scale), fl, chunk_size=1)
    run("C1 GBM 2^24 traj, final only", workloads.gbm(f64=True), scheme, 2 ** -8, int(2 ** 24 * scale), fl, chunk_size=None)
for scheme, fl in (("euler", 150), ("milstein", 2500), ("wagner_platen", 10900)):
    run("C2 polar walk (drifted) 1e7 traj, dt=2^-7, final only", workloads.polar_walk(4.0, (2.0, 0.0), 1.0, f64=True),
        scheme, 2 ** -7, int(1e7 * scale), fl, chunk_size=None,
        observables=workloads.polar_error_observables(4.0, 2.0))
run("C3 two RPY spheres 1e6 traj, 500 steps, final only", workloads.two_spheres(tmax=5.0, f64=True), "euler", 0.01,
    int(1e6 * scale), 650, chunk_size=None)
run("four RPY beads (jfm) 1e6 traj, 400 steps, final only", workloads.four_beads(tmax=20.0, f64=True), "euler", 0.05,
    int(1e6 * scale), 2500, chunk_size=None)
run("C5 GARCH call payoff 1e8 paths, 200 steps, moments only", workloads.garch(f64=True), "wagner_platen", 0.01,
    int(1e8 * scale), 10900, chunk_size=None, observables=workloads.call_payoff(120.0), store_trajectories=False)
json.dump({"fp64_peak_tflops": peak, "rows": rows}, open("gpurun_out/configs_r1.json", "w"), indent=1)
