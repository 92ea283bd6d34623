"""Round-2 exploratory measurements (run on the GPU box): numbers that the new tests' tolerances are set from."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import stochastic_b200 as pychastic
import stochastic_b200.numpy as jnp
from stochastic_b200 import workloads, sde_solver

out = {}
what = set(sys.argv[1:]) or {"hit", "sinh", "snap", "msd", "rot", "host"}
# "msdlong": the long-time diffusion coefficient of the four-bead chain (reference examples/jfm.py:221: 0.2898)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


if "hit" in what:
    import test_gpu_reference_data as T
    G = os.path.join(ROOT, "tests", "golden")
    strong = np.loadtxt(os.path.join(G, "strong_convergence_test.csv"), delimiter=",")
    weak = np.loadtxt(os.path.join(G, "weak_convergence_test.csv"), delimiter=",")
    rows = []
    for scheme, col in (("euler", 1), ("milstein", 3), ("wagner_platen", 5)):
        for n_steps in T.N_STEPS:
            row = len(T.N_STEPS) - 1 - T.N_STEPS.index(n_steps)
            r1 = T.hitting_time_errors(scheme, n_steps, 10_000)
            r = r1
            rows.append(dict(scheme=scheme, n_steps=n_steps, ours=r, ours_10k_seed0=r1, csv_strong=strong[row, col],
                             csv_strong_sigma=strong[row, col + 1], csv_weak=weak[row, col], csv_weak_sigma=weak[row, col + 1]))
            print(scheme, n_steps, "strong %.5f (10k: %.5f) csv %.5f | weak %.5f (10k: %.5f) +- %.5f csv %.5f" % (
                r["rmse"], r1["rmse"], strong[row, col], abs(r["mean"] - 0.5), abs(r1["mean"] - 0.5), r["mean_std"], weak[row, col]),
                "| 10k: mean_err %.5f err_std %.5f err_std/sqrt(n) %.6f mean_real-0.5 %.5f csv_sigmas %.5f %.6f" % (
                r1["mean_err"], r1["err_std"], r1["err_std"] / np.sqrt(r1["n"]), r1["mean_real"] - 0.5, strong[row, col + 1], weak[row, col + 1]), flush=True)
    out["hit"] = rows

if "fp" in what:
    Y_DRIFT, Y_BARRIER = 4.0, 2.0
    n = 2_000_000
    for scheme in ("euler", "milstein", "wagner_platen"):
        for n_steps in (64, 256):
            prob = workloads.polar_walk(Y_DRIFT, (2.0, 0.0), 2.0, f64=True)
            solver = pychastic.sde_solver.SDESolver(dt=prob.tmax / n_steps, scheme=scheme)
            kw = dict(n_trajectories=n, seed=0, chunk_size=None, store_trajectories=False)
            sol = solver.solve_many(prob, first_passage=lambda x, w, t: x[0] * jnp.sin(x[1]) - Y_BARRIER, **kw)
            real = solver.solve_many(prob, first_passage=lambda x, w, t: w[1] + Y_DRIFT * t - Y_BARRIER, **kw)
            both = sol["hit_flag"].torch() & real["hit_flag"].torch()
            err = (sol["hit_time"].torch() - real["hit_time"].torch())[both]
            print("fp", scheme, n_steps, "rmse", float(torch.sqrt((err ** 2).mean())), "both", float(both.float().mean()),
                  "mean real", float(real["hit_time"].torch()[both].mean()), flush=True)

if "sinh" in what:
    import problems as P
    res = []
    for scheme in ("euler", "milstein", "wagner_platen"):
        for e in range(3, 10):
            prob, _ = P.sinh_drift(1.0)
            P.as_f64(prob)
            n = 1 << 21
            s = pychastic.sde_solver.SDESolver(scheme=scheme, dt=2.0 ** -e)
            sol = s.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None, store_trajectories=False,
                               observables=lambda x, w, t: jnp.array([x[0], jnp.sinh(w[0] + t)]))
            mom = np.asarray(sol["moments"])
            res.append((scheme, e, mom[0, 0] / n - 1.937579205326131, mom[1, 0] / n - 1.937579205326131,
                        mom[0, 1] / n - (mom[0, 0] / n) ** 2))
            print("sinh", res[-1], flush=True)
    out["sinh"] = res

if "snap" in what:
    prob = workloads.polar_walk(4.0, (2.0, 0.0), 1.0, f64=True)
    res = {}
    for n, e in ((1 << 20, 7), (1 << 21, 6)):
        s = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=2.0 ** -e)
        t_final = timed(lambda: s.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None))
        t_snap = timed(lambda: s.solve_many(prob, n_trajectories=n, seed=0, chunk_size=1))
        s.warp_specialize = False
        t_snap_single = timed(lambda: s.solve_many(prob, n_trajectories=n, seed=0, chunk_size=1))
        t_final_single = timed(lambda: s.solve_many(prob, n_trajectories=n, seed=0, chunk_size=None))
        res[f"{n}x{2 ** e}"] = dict(ws_final=t_final, ws_every_step=t_snap, single_final=t_final_single,
                                   single_every_step=t_snap_single, bytes=n * 2 ** e * 40)
        print("snap", n, e, res[f"{n}x{2 ** e}"], flush=True)
    gb = workloads.gbm(f64=True)
    for scheme in ("euler", "milstein", "wagner_platen"):
        s = pychastic.sde_solver.SDESolver(scheme=scheme, dt=2.0 ** -8)
        n = 1 << 21
        t_snap = timed(lambda: s.solve_many(gb, n_trajectories=n, seed=0, chunk_size=1))
        t_final = timed(lambda: s.solve_many(gb, n_trajectories=n, seed=0, chunk_size=None))
        res[f"gbm_{scheme}"] = dict(every_step=t_snap, final=t_final, store_gbs=n * 512 * 24 / t_snap / 1e9)
        print("snap gbm", scheme, res[f"gbm_{scheme}"], flush=True)
    out["snap"] = res

if "msd" in what:
    prob = workloads.four_beads(tmax=400.0, f64=True)
    radii = np.array([3.0, 1.0, 1.0, 1.0])
    from stochastic_b200 import grpy

    def trace_mu(x, w, t):
        mu = grpy.muTT(jnp.reshape(x, (4, 3)), radii)
        return jnp.array([mu[3 * a + 0][3 * b + 0] + mu[3 * a + 1][3 * b + 1] + mu[3 * a + 2][3 * b + 2]
                          for a in range(4) for b in range(4)])
    s = pychastic.sde_solver.SDESolver(dt=0.05)
    n = 1 << 16
    t0 = time.perf_counter()
    sol = s.solve_many(prob, n_trajectories=n, chunk_size=40, chunks_per_randomization=1, seed=0,
                       store_trajectories=False, observables=trace_mu)
    tm = (np.asarray(sol["moments"])[:, 0] / n).reshape(4, 4)
    inv = np.linalg.inv(tm)
    weights = inv.sum(-1) / inv.sum()
    mstdc = 6.0 / inv.sum()
    print("weights", weights, "mstdc", mstdc, time.perf_counter() - t0, flush=True)
    wts = [float(v) for v in weights]

    def msd(x, w, t, xr):
        d = x - xr
        c = [sum(wts[a] * d[3 * a + k] for a in range(4)) for k in range(3)]
        return jnp.array([d[0] ** 2 + d[1] ** 2 + d[2] ** 2, d[9] ** 2 + d[10] ** 2 + d[11] ** 2,
                          c[0] ** 2 + c[1] ** 2 + c[2] ** 2])
    t0 = time.perf_counter()
    sol = s.solve_many(prob, n_trajectories=n, chunk_size=40, chunks_per_randomization=1, seed=0,
                       store_trajectories=False, observables=msd, observe_at="snapshots")
    mom = np.asarray(sol["moments"])
    print("msd pass", time.perf_counter() - t0, mom.shape, flush=True)
    tt = 2.0 * np.arange(mom.shape[0])
    m = mom[:, :, 0] / n
    G = os.path.join(ROOT, "tests", "golden")
    for k, name in enumerate(("big_bead", "small_bead", "centre")):
        ref = np.loadtxt(os.path.join(G, f"bead_{name}_msd_head.csv"))
        K = min(len(ref), len(m))
        rel = np.abs(m[1:K, k] - ref[1:K]) / ref[1:K]
        a = np.polyfit(tt[20:], m[20:, k], 1)[0]
        print(name, "slope*pi", a * np.pi, "max rel dev vs reference csv (256 traj)", rel.max(), "first rows", m[:4, k], ref[:4], flush=True)
    out["msd"] = dict(weights=wts, mstdc=float(mstdc), msd=m.tolist())

if "msdlong" in what:
    from stochastic_b200 import grpy
    radii = np.array([3.0, 1.0, 1.0, 1.0])
    for half in (True, False):
        prob = workloads.four_beads(tmax=8000.0, half_spring=half, f64=True)

        def trace_mu(x, w, t):
            mu = grpy.muTT(jnp.reshape(x, (4, 3)), radii)
            return jnp.array([mu[3 * a + 0][3 * b + 0] + mu[3 * a + 1][3 * b + 1] + mu[3 * a + 2][3 * b + 2]
                              for a in range(4) for b in range(4)])
        s = pychastic.sde_solver.SDESolver(dt=0.05)
        n = 1 << 14
        t0 = time.perf_counter()
        sol = s.solve_many(prob, n_trajectories=n, chunk_size=40, chunks_per_randomization=1, seed=0,
                           store_trajectories=False, observables=trace_mu)
        tm = (np.asarray(sol["moments"])[:, 0] / n).reshape(4, 4)
        inv = np.linalg.inv(tm)
        wts = [float(v) for v in inv.sum(-1) / inv.sum()]
        print("half", half, "weights", wts, "mstdc", 6.0 / inv.sum(), time.perf_counter() - t0, flush=True)

        def msd(x, w, t, xr):
            d = x - xr
            c = [sum(wts[a] * d[3 * a + k] for a in range(4)) for k in range(3)]
            return jnp.array([d[0] ** 2 + d[1] ** 2 + d[2] ** 2, d[9] ** 2 + d[10] ** 2 + d[11] ** 2,
                              c[0] ** 2 + c[1] ** 2 + c[2] ** 2])
        sol = s.solve_many(prob, n_trajectories=n, chunk_size=40, chunks_per_randomization=1, seed=0,
                           store_trajectories=False, observables=msd, observe_at="snapshots")
        m = np.asarray(sol["moments"])[:, :, 0] / n
        tt = 2.0 * np.arange(m.shape[0])
        for lo, hi in ((20, 200), (200, 1000), (1000, 2000), (2000, 4000), (500, 4000)):
            print("half", half, "lags", lo * 2, hi * 2, "slope*pi big/small/centre",
                  [float(np.polyfit(tt[lo:hi], m[lo:hi, k], 1)[0] * np.pi) for k in range(3)], flush=True)
        out["msdlong_%s" % half] = m[::20].tolist()

if "rot" in what:
    p, post, obs = workloads.evensen_rotations(f64=False)
    s = pychastic.sde_solver.SDESolver(dt=0.008)
    n = 1 << 18
    sol = s.solve_many(p, n_trajectories=n, step_post_processing=post, chunks_per_randomization=2, seed=0,
                       store_trajectories=False, observables=obs, observe_at="snapshots")
    mom = np.asarray(sol["moments"])[:, :, 0] / n
    t = 0.008 * np.arange(1, mom.shape[0] + 1)
    th = 1 / 6 - 5 / 12 * np.exp(-6 * t) + 1 / 4 * np.exp(-2 * t)
    G = os.path.join(ROOT, "tests", "golden")
    for k, c in enumerate((1, 3, 5)):
        ref = np.loadtxt(os.path.join(G, f"rotational_curve_{c}.csv"), delimiter=",")
        print("rot comp", k, "max |ours - theory|", np.abs(mom[:250, k] - th[:250]).max(), "max |csv - theory|",
              np.abs(ref[:, 1] - th[:250]).max(), "max |ours - csv|", np.abs(mom[:250, k] - ref[:, 1]).max(), flush=True)
    out["rot"] = dict(ours=mom[:250].tolist())

if "host" in what:
    res = {}
    for name, prob, scheme in (("polar_wp", workloads.polar_walk(4.0, (2.0, 0.0), f64=True), "wagner_platen"),
                               ("two_spheres", workloads.two_spheres(tmax=0.1, f64=True), "euler"),
                               ("four_beads", workloads.four_beads(tmax=0.5, f64=True), "euler")):
        s = pychastic.sde_solver.SDESolver(scheme=scheme, dt=0.01)
        s.solve_many(prob, n_trajectories=16, chunk_size=None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            s.solve_many(prob, n_trajectories=16, chunk_size=None)
        torch.cuda.synchronize()
        res[name] = (time.perf_counter() - t0) / 200
        print("host", name, "%.3f ms per call" % (1e3 * res[name]), sde_solver.PLAN_STATS, flush=True)
    out["host"] = res

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "explore_r2_" + "_".join(sorted(what)) + ".json"), "w") as f:
    json.dump(out, f, default=float)
