"""Snapshot-heavy runs (chunk_size = 1: every step stored): useful store bandwidth with the TMA tile path
(csrc/sde_snap_tma.cuh), the transposed write-out loop and direct stores, next to the final-only time.
python profiles/run_snapshots.py [scale]  ->  gpurun_out/snapshots_r2.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stochastic_b200 as pychastic
from stochastic_b200 import workloads

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
HBM_PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6554.0) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6554.0
rows = []


def run(name, prob, scheme, dt, n, mode, reps=3, ws=None, **kw):
    solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=dt)
    solver.warp_specialize = ws
    old = pychastic.config.get("snapshot_tma")
    if mode == "loop":
        pychastic.config.update("snapshot_tma", False)
    elif mode == "direct":
        pychastic.config.update("snapshot_tma", False)
        solver.snap_tile = 1
    elif isinstance(mode, tuple):
        solver.snap_tile = mode
    best = 1e30
    try:
        for _ in range(reps):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            sol = solver.solve_many(prob, n_trajectories=n, seed=0, progress_bar=False, **kw)
            ev1.record()
            torch.cuda.synchronize()
            best = min(best, ev0.elapsed_time(ev1))
            del sol
    finally:
        pychastic.config.update("snapshot_tma", old)
    steps = solver.last_steps
    rate = n * steps / best * 1e3
    row = {"config": name, "scheme": scheme, "mode": str(mode), "n_traj": n, "steps": steps, "ms": round(best, 3),
           "traj_steps_per_s": rate, "registers": solver.last_program.info()["registers"]}
    if kw.get("chunk_size", 1) == 1:
        d, m = prob.dimension, prob.noise_terms
        row["store_GBps"] = rate * (1 + d + m) * 8 / 1e9
        row["frac_hbm_peak"] = row["store_GBps"] / HBM_PEAK
    rows.append(row)
    print(json.dumps(row), flush=True)


n = int(2 ** 21 * scale)
for scheme in ("euler", "milstein", "wagner_platen"):
    for mode in ("auto", ("tma", 8, 2), ("tma", 4, 1), "loop", "direct"):
        run("C1 GBM 2^21 x 512, every step stored", workloads.gbm(f64=True), scheme, 2 ** -8, n, mode, chunk_size=1)
    run("C1 GBM 2^21 x 512, final only", workloads.gbm(f64=True), scheme, 2 ** -8, n, "final", chunk_size=None)
n2 = int(2 ** 20 * scale)
for ws in (True, False):
    for mode in ("auto", "loop", "direct"):
        run(f"C2 polar walk 2^20 x 128, every step stored, ws={ws}", workloads.polar_walk(4.0, (2.0, 0.0), 1.0, f64=True),
            "wagner_platen", 2 ** -7, n2, mode, ws=ws, chunk_size=1)
    run(f"C2 polar walk 2^20 x 128, final only, ws={ws}", workloads.polar_walk(4.0, (2.0, 0.0), 1.0, f64=True),
        "wagner_platen", 2 ** -7, n2, "final", ws=ws, chunk_size=None)
for scheme in ("euler", "milstein"):
    for mode in ("auto", "loop"):
        run("C2 polar walk 2^20 x 128, every step stored", workloads.polar_walk(4.0, (2.0, 0.0), 1.0, f64=True), scheme,
            2 ** -7, n2, mode, chunk_size=1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"hbm_peak_GBps": HBM_PEAK, "rows": rows}, open("gpurun_out/snapshots_r2.json", "w"), indent=1)
