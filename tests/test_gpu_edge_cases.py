"""GPU parity, part 2: larger noise dimension, the partitionable PRNG layout, ragged sizes for the staged
snapshot stores, prefixes of long runs, empty runs, device-side moments and the result-array surface."""
import numpy as np
import pytest
import torch

import stochastic_b200 as pychastic
import stochastic_b200.numpy as jnp
from stochastic_b200 import workloads
from oracle import solver as osolver

import problems as P

pytestmark = pytest.mark.gpu


def _close64(got, ref, rtol=1e-9):
    for k in ("time_values", "solution_values", "wiener_values"):
        g, r = np.asarray(got[k]), ref[k]
        assert g.shape == r.shape, (k, g.shape, r.shape)
        err = np.abs(g - r) / np.maximum(np.abs(r), 1e-3)
        assert err.max() < rtol, (k, err.max())


@pytest.mark.parametrize("scheme", ["euler", "milstein", "wagner_platen"])
@pytest.mark.parametrize("name", ["coupled_3x3", "rectangular_2x3"])
def test_three_noise_terms(name, scheme):
    prob, oprob = getattr(P, name)()
    P.as_f64(prob, oprob)
    solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=prob.tmax / 8)
    got = solver.solve_many(prob, n_trajectories=9, seed=11, chunk_size=2)
    ref = osolver.solve_many(oprob, scheme=scheme, dt=prob.tmax / 8, n_trajectories=9, seed=11, chunk_size=2,
                             dtype=np.float64)
    _close64(got, ref)


@pytest.mark.parametrize("scheme", ["euler", "milstein", "wagner_platen"])
def test_partitionable_layout(scheme):
    """jax >= 0.5 default key / bit layout (jax_threefry_partitionable=True)."""
    prob, oprob = P.polar_walk(4.0, (2.0, 0.0), 0.25)
    P.as_f64(prob, oprob)
    solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=2 ** -5)
    got = solver.solve_many(prob, n_trajectories=10, seed=6, chunks_per_randomization=2, prng_layout="partitionable")
    ref = osolver.solve_many(oprob, scheme=scheme, dt=2 ** -5, n_trajectories=10, seed=6, chunks_per_randomization=2,
                             dtype=np.float64, layout="partitionable")
    _close64(got, ref)
    other = solver.solve_many(prob, n_trajectories=10, seed=6, chunks_per_randomization=2)
    assert not np.allclose(np.asarray(other["wiener_values"]), ref["wiener_values"])


@pytest.mark.parametrize("n_traj,steps,chunk", [(45, 23, 1), (33, 37, 2), (1, 16, 1), (70, 5, 1), (64, 40, 3)])
def test_ragged_sizes_through_the_staged_snapshot_stores(n_traj, steps, chunk):
    """Trajectory counts that are not multiples of the warp size and snapshot counts that are not multiples of the
    staging tile (SDE_SNAP_TILE = 16)."""
    prob, oprob = P.gbm(0.2, 0.5, 1.0, 1.0)
    P.as_f64(prob, oprob)
    for scheme in ("euler", "wagner_platen"):
        solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=1.0 / steps)
        got = solver.solve_many(prob, n_trajectories=n_traj, seed=2, chunk_size=chunk)
        ref = osolver.solve_many(oprob, scheme=scheme, dt=1.0 / steps, n_trajectories=n_traj, seed=2, chunk_size=chunk,
                                 dtype=np.float64, use_closed_form=True)
        _close64(got, ref)


def test_staged_stores_for_vector_state_and_multiple_ics():
    prob, oprob = P.polar_walk(0.0, (1.0, 0.0), 1.0)
    x0 = np.array([[1.0, 0.0], [2.0, 0.5], [1.5, -0.3], [3.0, 1.0], [0.8, 2.0]])
    prob.x0 = x0
    oprob.x0 = x0
    solver = pychastic.sde_solver.SDESolver(scheme="milstein", dt=1 / 40)
    got = solver.solve_many(prob, n_trajectories=None, seed=9)
    ref = osolver.solve_many(oprob, scheme="milstein", dt=1 / 40, n_trajectories=None, seed=9, dtype=np.float64)
    assert np.asarray(got["solution_values"]).shape == (5, 40, 2)
    _close64(got, ref)


def test_max_snapshots_is_a_prefix_of_the_full_run():
    prob, _ = P.polar_walk(4.0, (2.0, 0.0), 1.0)
    P.as_f64(prob)
    solver = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=1 / 64)
    full = solver.solve_many(prob, n_trajectories=20, seed=1, chunk_size=4, chunks_per_randomization=2)
    head = solver.solve_many(prob, n_trajectories=20, seed=1, chunk_size=4, chunks_per_randomization=2, max_snapshots=5)
    for k in ("time_values", "solution_values", "wiener_values"):
        assert np.array_equal(np.asarray(full[k])[:, :5], np.asarray(head[k]))


def test_empty_run():
    prob, _ = P.gbm()
    sol = pychastic.sde_solver.SDESolver(dt=0.25).solve_many(prob, n_trajectories=0)
    assert np.asarray(sol["solution_values"]).shape == (0, 8, 1)


def test_call_payoff_moments_match_oracle():
    """BASELINE config 5 in miniature: GARCH model, call payoff reduced on the device."""
    prob, oprob = P.garch(tmax=0.5)
    P.as_f64(prob, oprob)
    solver = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=0.05)
    got = solver.solve_many(prob, n_trajectories=300, seed=4, chunk_size=None, observables=workloads.call_payoff(100.0),
                            store_trajectories=False)
    assert "solution_values" not in got
    ref = osolver.solve_many(oprob, scheme="wagner_platen", dt=0.05, n_trajectories=300, seed=4, chunk_size=None,
                             dtype=np.float64)
    payoff = np.maximum(ref["solution_values"][:, -1, 0] - 100.0, 0.0)
    mom = np.asarray(got["moments"])
    assert got["count"] == 300
    assert np.isclose(mom[0, 0], payoff.sum(), rtol=1e-9) and np.isclose(mom[0, 1], (payoff ** 2).sum(), rtol=1e-9)


def test_result_arrays_behave_like_numpy_and_export_dlpack():
    prob, _ = P.gbm()
    sol = pychastic.sde_solver.SDESolver(dt=0.125).solve_many(prob, n_trajectories=6, seed=0)
    x = sol["solution_values"]
    x.block_until_ready()
    assert x.shape == (6, 16, 1) and x.ndim == 3 and x.dtype == np.float32 and len(x) == 6
    host = np.asarray(x)
    assert np.array_equal(np.asarray(x[:, -1]), host[:, -1])
    assert np.array_equal(np.asarray(x[2:4, ::2, 0]), host[2:4, ::2, 0])
    assert np.allclose(np.asarray(x.reshape(-1, 1)), host.reshape(-1, 1))
    assert np.isclose(float(jnp.mean(x)), host.mean()) and np.isclose(float(x.mean()), host.mean())
    assert np.allclose(np.asarray(2.0 * x + 1.0), 2.0 * host + 1.0)
    t = torch.from_dlpack(x)
    assert t.is_cuda and t.data_ptr() == x.torch().data_ptr()
    (r,) = np.asarray(x[:, -1, :]).transpose()
    assert r.shape == (6,)


def test_brownian_rotations_with_step_post_processing():
    """the SO(3) rotational-diffusion example (lax.cond branches, jacobian inside the drift,
    canonicalisation hook after every step), reference examples/examples_from_paper/brownian_rotations.py."""
    from oracle import problems as oprob
    prob, post = workloads.brownian_rotations(tmax=2.0, f64=True)
    oracle_prob, opost = oprob.brownian_rotations(tmax=2.0)
    solver = pychastic.sde_solver.SDESolver(dt=0.01)
    got = solver.solve_many(prob, step_post_processing=post, n_trajectories=24, chunk_size=50,
                            chunks_per_randomization=1, seed=3)
    ref = osolver.solve_many(oracle_prob, scheme="euler", dt=0.01, n_trajectories=24, chunk_size=50,
                             chunks_per_randomization=1, seed=3, dtype=np.float64, step_post_processing=opost)
    _close64(got, ref, rtol=1e-7)          # 200 chaotic steps near the coordinate singularity amplify rounding
    angles = np.sqrt((np.asarray(got["solution_values"]) ** 2).sum(-1))
    assert angles.max() <= np.pi + 1e-9


@pytest.mark.parametrize("scheme", ["euler", "wagner_platen"])
def test_first_passage_observer_matches_host_post_processing(scheme):
    """the hitting-time study of reference examples/hit.py (barrier y = r sin(phi) > 2) done on
    the device, against the reference's own host-side process_trajectory logic applied to the stored trajectory."""
    y_barrier = 2.0
    prob, oprob = P.polar_walk(4.0, (2.0, 0.0), 1.0)
    P.as_f64(prob, oprob)
    dt = 2 ** -6
    solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=dt)
    event = lambda x, w, t: x[0] * jnp.sin(x[1]) - y_barrier
    sol = solver.solve_many(prob, n_trajectories=200, seed=0, chunks_per_randomization=4, first_passage=event)
    traj = np.asarray(sol["solution_values"])
    tvals = np.asarray(sol["time_values"])
    flag, ht, hx = (np.asarray(sol[k]) for k in ("hit_flag", "hit_time", "hit_state"))
    y = traj[:, :, 0] * np.sin(traj[:, :, 1])
    n_hit = 0
    for i in range(200):
        above = y[i] > y_barrier
        assert bool(flag[i]) == bool(above.any())
        if not above.any():
            continue
        k = int(above.argmax())
        before = traj[i, k - 1] if k > 0 else np.asarray(prob.x0)
        yb = y[i, k - 1] if k > 0 else prob.x0[0] * np.sin(prob.x0[1])
        wgt = (y_barrier - yb) / (y[i, k] - yb)
        assert np.allclose(hx[i], traj[i, k] * wgt + before * (1 - wgt), rtol=1e-9, atol=1e-12)
        assert np.isclose(ht[i], (tvals[i, k] - dt) + wgt * dt, rtol=1e-9)
        n_hit += 1
    assert n_hit > 50          # drift 4 towards a barrier at distance 2: most paths hit within T = 1


@pytest.mark.parametrize("n_steps", [1, 2, 3])
@pytest.mark.parametrize("warp_specialize", [True, False])
def test_very_short_runs_in_both_stepping_kernels(n_steps, warp_specialize):
    """One, two and three steps in total: the producer/consumer ring of the warp-specialised kernel has two stages,
    so these are exactly the cases where its empty/full hand-shake starts and ends."""
    prob, oprob = P.polar_walk(4.0, (2.0, 0.0), 0.25)
    P.as_f64(prob, oprob)
    dt = 0.25 / n_steps
    solver = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=dt)
    solver.warp_specialize = warp_specialize
    got = solver.solve_many(prob, n_trajectories=130, seed=8, chunk_size=None)     # 130: one full CTA + 2 trajectories
    ref = osolver.solve_many(oprob, scheme="wagner_platen", dt=dt, n_trajectories=130, seed=8, chunk_size=None,
                             dtype=np.float64)
    _close64(got, ref)


def test_both_stepping_kernels_agree_bit_for_bit():
    """Single-role kernel, warp-specialised kernel, and the warp-specialised kernel whose producers draw uniforms and
    whose consumers apply erf_inv (SDE_WS_SPLIT_NORMAL, a measured-and-rejected variant kept as a tuning knob)."""
    prob, _ = P.polar_walk(4.0, (2.0, 0.0), 1.0)
    P.as_f64(prob)
    out = []
    for ws, tuning in ((True, None), (False, None), (True, {"SDE_WS_SPLIT_NORMAL": 1})):
        solver = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=2 ** -5)
        solver.warp_specialize = ws
        solver.ws_tuning = tuning
        out.append(solver.solve_many(prob, n_trajectories=777, seed=5, chunk_size=4, chunks_per_randomization=2))
    for other in out[1:]:
        for k in ("time_values", "solution_values", "wiener_values"):
            assert np.array_equal(np.asarray(out[0][k]), np.asarray(other[k]))


def _stores_three_ways(prob, scheme, dt, n_traj, tma_choices, **kw):
    """The same run with snapshots stored (a) by the TMA unit, (b) by the transposed write-out loop, (c) directly."""
    out = []
    for choice in list(tma_choices) + ["loop", "direct"]:
        solver = pychastic.sde_solver.SDESolver(scheme=scheme, dt=dt)
        solver.warp_specialize = kw.get("warp_specialize")
        old = pychastic.config.get("snapshot_tma")
        try:
            if choice == "loop":
                pychastic.config.update("snapshot_tma", False)
            elif choice == "direct":
                pychastic.config.update("snapshot_tma", False)
                solver.snap_tile = 1
            elif choice != "auto":
                solver.snap_tile = choice
            args = {k: v for k, v in kw.items() if k != "warp_specialize"}
            out.append((choice, solver.solve_many(prob, n_trajectories=n_traj, seed=4, **args)))
        finally:
            pychastic.config.update("snapshot_tma", old)
    ref = out[-1][1]
    for choice, sol in out[:-1]:
        for k in ("time_values", "solution_values", "wiener_values"):
            assert np.array_equal(np.asarray(sol[k]), np.asarray(ref[k])), (choice, k)
    return ref


@pytest.mark.parametrize("n_traj,steps", [(45, 24), (33, 38), (1, 16), (70, 6), (64, 40), (300, 130), (129, 37)])
def test_tma_snapshot_tiles_equal_plain_stores_scalar_f64(n_traj, steps):
    """Snapshot tiles handed to the TMA unit (csrc/sde_snap_tma.cuh), every row length and swizzle pattern, against the
    write-out loop and the direct stores, bit for bit: ragged trajectory counts (rows past the end are clipped by the
    tensor map), runs that end inside a tile, an odd snapshot count (rows no tensor map can describe: falls back)."""
    prob, _ = P.gbm(0.2, 0.5, 1.0, 1.0)
    P.as_f64(prob)
    choices = ["auto"]
    if steps % 2 == 0:
        choices += [("tma", 16, 3), ("tma", 8, 2), ("tma", 4, 1), ("tma", 2, 0)]
    _stores_three_ways(prob, "euler", 1.0 / steps, n_traj, choices)


@pytest.mark.parametrize("scheme,ws", [("milstein", None), ("wagner_platen", False), ("wagner_platen", True)])
def test_tma_snapshot_tiles_equal_plain_stores_vector_state(scheme, ws):
    """d = m = 2: the tile of a field of width 2 is two boxes; also the consumers of the warp-specialised kernel."""
    prob, _ = P.polar_walk(4.0, (2.0, 0.0), 1.0)
    P.as_f64(prob)
    # (the ring of the warp-specialised kernel leaves room for 16-byte rows only; the single-role Wagner-Platen kernel
    # stages 94 KB of normals per 256-thread CTA, which rules out 128-byte rows)
    choices = ["auto", ("tma", 2, 0)] + ([] if ws else [("tma", 8, 2), ("tma", 4, 1)]) + ([("tma", 16, 3)] if ws is None else [])
    _stores_three_ways(prob, scheme, 1.0 / 44, 200, choices, warp_specialize=ws)
    _stores_three_ways(prob, scheme, 1.0 / 64, 131, choices, warp_specialize=ws, chunk_size=2, chunks_per_randomization=4)


def test_tma_snapshot_tiles_equal_plain_stores_f32():
    prob, _ = P.gbm(0.2, 0.5, 1.0, 1.0)
    _stores_three_ways(prob, "euler", 1.0 / 72, 77, ["auto", ("tma", 32, 3), ("tma", 16, 2), ("tma", 8, 1), ("tma", 4, 0)])
    prob, _ = P.polar_walk(4.0, (2.0, 0.0), 1.0)
    _stores_three_ways(prob, "euler", 1.0 / 36, 50, ["auto", ("tma", 32, 3), ("tma", 4, 0)])
