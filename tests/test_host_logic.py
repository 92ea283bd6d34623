"""CPU tests of the host side: tracer, symbolic Taylor-Ito tensors, emitter, problem validation,
chunk arithmetic and the analytic moments module.  No GPU, no compute calls into the CUDA library."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

import stochastic_b200 as pychastic
import stochastic_b200.numpy as jnp
from stochastic_b200 import grpy, workloads
from stochastic_b200.sde_solver import L_t_operator, L_w_operator, build_source, chunk_arithmetic
from stochastic_b200.trace import emit
from stochastic_b200.trace import expr as E
from stochastic_b200.trace.taylor import TracedProblem, step_expressions
from oracle import problems as oprob, rpy as orpy, solver as osolver, taylor_ito

import problems as P


# -- SDEProblem ---------------------------------------------------------------------------------
def test_problem_scalar_and_vector_inputs():
    """reference tests/test_sde_problem.py"""
    p = pychastic.sde_problem.SDEProblem(lambda x: jnp.array(0), lambda x: jnp.array(0), 0, 1)
    assert (p.dimension, p.noise_terms, p.x0.shape) == (1, 1, (1,))
    p = pychastic.sde_problem.SDEProblem(lambda x: jnp.zeros(2), lambda x: jnp.zeros((2, 3)), jnp.zeros(2), 1)
    assert (p.dimension, p.noise_terms) == (2, 3)
    assert p.x0.dtype == np.float32          # reference sde_problem.py:51


def test_problem_validation_errors():
    with pytest.raises(ValueError):
        pychastic.sde_problem.SDEProblem(lambda x: x, lambda x: x, 1.0, tmax=0)
    with pytest.raises(ValueError):
        pychastic.sde_problem.SDEProblem(lambda x: jnp.zeros(3), lambda x: jnp.zeros((2, 2)), jnp.zeros(2), 1)
    with pytest.raises(ValueError):
        pychastic.sde_problem.SDEProblem(lambda x: jnp.zeros(2), lambda x: jnp.zeros(2), jnp.zeros(2), 1)
    p = pychastic.sde_problem.SDEProblem(lambda x: jnp.zeros(2), lambda x: jnp.zeros((2, 2)), jnp.zeros((5, 2)), 1)
    assert p.x0.shape == (5, 2) and p.dimension == 2


def test_x64_switch():
    pychastic.config.update("enable_x64", True)
    try:
        p = pychastic.sde_problem.SDEProblem(lambda x: x, lambda x: x, 0.1, 1.0)
        assert p.x0.dtype == np.float64 and p.x0[0] == 0.1
    finally:
        pychastic.config.update("enable_x64", False)


# -- host arithmetic ------------------------------------------------------------------------------
@pytest.mark.parametrize("tmax,dt,chunk,cpr", [(2.0, 2 ** -8, 1, None), (1.0, 1 / 23, 3, 2), (1.0, 0.01, None, None),
                                               (8000.0, 0.05, 40, 1), (1.0, 0.3, 1, None), (2.0, 2 ** -10, 1, 3)])
def test_chunk_arithmetic_matches_oracle(tmax, dt, chunk, cpr):
    assert chunk_arithmetic(tmax, dt, chunk, cpr) == osolver.chunk_arithmetic(tmax, dt, chunk, cpr)


def test_chunk_arithmetic_overshoot():
    chunk, n_chunks, cpr, n_frag = chunk_arithmetic(1.0, 1 / 23, 3, 2)
    assert (chunk, n_chunks, cpr, n_frag) == (3, 8, 2, 4)      # 24 steps integrated for 23 needed


def test_unknown_scheme_raises_before_touching_the_gpu():
    prob = workloads.gbm()
    with pytest.raises(ValueError):
        pychastic.sde_solver.SDESolver(scheme="heun").solve_many(prob, 4)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        pychastic.sde_solver.SDESolver().solve_many(workloads.gbm(), 4)


# -- L operators (the reference's own known-answer tests) -------------------------------------------
def _L(f, idx, problem):
    for c in reversed(idx):
        f = L_t_operator(f, problem) if c == "t" else L_w_operator(f, problem)
    return f


def test_L_operator_logic_1d():
    """reference tests/test_coefficient_function_logic_1d.py:43-56"""
    problem = pychastic.sde_problem.SDEProblem(a=lambda x: jnp.exp(3.0 * x), b=lambda x: jnp.exp(7.0 * x), tmax=1.0,
                                               x0=jnp.array(0.0))
    x0 = np.array([0.0])
    ident = lambda x: x
    got = [np.squeeze(_L(ident, k, problem)(x0)) for k in ("t", "w", "ww", "tw", "wt")]
    assert np.allclose(got, [1.0, 1.0, 7.0, 7.0 + 0.5 * 7.0 * 7.0, 3.0])


def test_L_operator_logic_2d():
    """reference tests/test_coefficient_function_logic_2d.py:48-67 (index order of f_ww, f_wt)"""
    problem, _ = P.exp_2d()
    x0 = np.zeros(2)
    ident = lambda x: x
    f_t, f_w, f_ww, f_wt = (np.squeeze(_L(ident, k, problem)(x0)) for k in ("t", "w", "ww", "wt"))
    a = np.array([1.1, 1.2])
    b = np.array([[1.3, 1.4], [1.5, 1.6]])
    bp = np.array([[[1.3 * 7, 0], [0, 1.6 * 17]], [[0, 1.4 * 11], [1.5 * 13, 0]]])
    ap = np.array([[1.1 * 3, 0], [0, 1.2 * 5]])
    assert np.allclose(f_t, a) and np.allclose(f_w, b)
    assert np.allclose(f_ww, np.einsum("ai,abj->bij", b, bp))
    assert np.allclose(f_wt, np.einsum("cj,cd->dj", b, ap))


# -- symbolic coefficient tensors vs autodiff oracle ---------------------------------------------------
CASES = {
    "polar_drift": (lambda: P.polar_walk(4.0, (2.0, 0.0)), [1.7, 0.4]),
    "exp_2d": (P.exp_2d, [0.01, -0.02]),
    "garch": (P.garch, [90.0, 0.07]),
    "parabolic": (P.parabolic_walk, [1.1, 0.2]),
    "gbm": (P.gbm, [1.3]),
    "arctan": (P.arctan, [0.3]),
    "tanh": (P.tanh_benchmark, [0.4]),
    "sinh": (P.sinh_drift, [0.2]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_symbolic_coefficients_match_autodiff(name):
    mk, x0 = CASES[name]
    prob, op = mk()
    tp = TracedProblem(prob.a, prob.b, len(x0))
    cf = tp.coefficients("wagner_platen")
    ref = taylor_ito.coefficient_functions(op.a, op.b)
    xt = torch.tensor(x0, dtype=torch.float64)
    for k, v in cf.items():
        flat = [E.wrap(e) for e in np.array(v, dtype=object).ravel()]
        got = np.array([float(z) for z in E.evaluate(flat, np.array(x0))])
        want = ref[k](xt).numpy().ravel()
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12), k


def test_structural_zeros_are_dropped():
    """GARCH: b is diagonal, a_0 constant -> most Wagner-Platen tensors are structurally zero."""
    prob, _ = P.garch()
    tp = TracedProblem(prob.a, prob.b, 2)
    cf = tp.coefficients("wagner_platen")
    f_www = np.array(cf["f_www"], dtype=object)
    zeros = sum(1 for e in f_www.ravel() if e.is_const and e.value == 0.0)
    assert zeros >= 11 and f_www.size == 16


def _compile_c_step(prob, d, scheme):
    tp = TracedProblem(prob.a, prob.b, d)
    outs, sv = step_expressions(tp, scheme)
    names = {}
    for i in range(d):
        names[i] = f"x[{i}]"
    names[sv.dt] = "dt"
    for idx in range(d + 1, sv.count):
        names[idx] = f"I[{idx - d - 1}]"
    body = emit.emit_body([(f"xn[{i}]", e) for i, e in enumerate(outs)], names, f64=True, lang="c")
    src = ("#define _GNU_SOURCE\n#include <math.h>\n#include <stdbool.h>\ntypedef double real;\n"
           "void step(const real* x, const real* I, real dt, real* xn) {\n" + body + "\n}\n")
    tmp = tempfile.mkdtemp()
    c, so = os.path.join(tmp, "step.c"), os.path.join(tmp, "step.so")
    open(c, "w").write(src)
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", so, c, "-lm"])
    lib = ctypes.CDLL(so)
    lib.step.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_double, ctypes.c_void_p]
    return lib, sv


@pytest.mark.parametrize("name", ["polar_drift", "garch", "parabolic"])
def test_emitted_step_compiled_as_c_matches_oracle_step(name):
    """The emitter's text (the same body NVRTC compiles) run on the host against the oracle's contraction."""
    mk, x0 = CASES[name]
    prob, op = mk()
    d, m = 2, 2
    lib, sv = _compile_c_step(prob, d, "wagner_platen")
    rng = np.random.default_rng(5)
    dt = 0.01
    I = rng.normal(size=sv.count - d - 1) * 0.1
    x = np.array(x0, dtype=np.float64)
    xn = np.zeros(d)
    lib.step(x.ctypes.data, I.ctypes.data, dt, xn.ctypes.data)
    cf = {k: f(torch.tensor(x)).numpy() for k, f in taylor_ito.coefficient_functions(op.a, op.b).items()}
    off = 0
    dW = I[off:off + m]; off += m
    Iww = I[off:off + m * m].reshape(m, m); off += m * m
    Iwt = I[off:off + m]; off += m
    Itw = I[off:off + m]; off += m
    Iwww = I[off:off + m ** 3].reshape(m, m, m)
    want = (x + cf["f_t"].reshape(d) * dt + cf["f_w"] @ dW + np.einsum("ijk,jk->i", cf["f_ww"], Iww)
            + cf["f_tw"].reshape(d, m) @ Itw + cf["f_wt"].reshape(d, m) @ Iwt
            + np.einsum("ijkl,jkl->i", cf["f_www"], Iwww) + cf["f_tt"].reshape(d) * dt * dt / 2)
    assert np.allclose(xn, want, rtol=1e-12, atol=1e-13)


def test_generated_source_is_deterministic_and_cached_by_content():
    prob = workloads.polar_walk(4.0, (2.0, 0.0))
    s1 = build_source(prob.a, prob.b, 2, "wagner_platen", True, "original")[0]
    s2 = build_source(prob.a, prob.b, 2, "wagner_platen", True, "original")[0]
    assert s1 == s2 and "sincos" in s1 and "#define SDE_SCHEME 2" in s1


_EMIT_SCRIPT = """
import hashlib, sys
sys.path.insert(0, {root!r})
from stochastic_b200 import workloads
from stochastic_b200.sde_solver import build_source
def src(name):
    if name == "garch":
        p = workloads.garch(f64=True)
        return build_source(p.a, p.b, 2, "milstein", True, "original").source
    if name == "polar":
        p = workloads.polar_walk(4.0, (2.0, 0.0), f64=True)
        return build_source(p.a, p.b, 2, "wagner_platen", True, "original", observe=workloads.polar_error_observables()).source
    if name == "beads":
        p = workloads.four_beads(f64=True)
        return build_source(p.a, p.b, 12, "euler", True, "original").source
    p, post = workloads.brownian_rotations(f64=True)
    return build_source(p.a, p.b, 3, "euler", True, "original", post=post).source
for name in sys.argv[1:]:
    print(name, hashlib.sha256(src(name).encode()).hexdigest())
"""


def test_emitted_source_does_not_depend_on_trace_history(tmp_path):
    """The CUDA text of a problem (hence its cubin-cache key and its FMA contraction) must be a pure function of the
    problem: commutative operands are ordered by a structural key, not by the order in which a process happened to
    create the nodes, and not by Python's per-process string hashing."""
    script = tmp_path / "emit.py"
    script.write_text(_EMIT_SCRIPT.format(root=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    runs = []
    for seed, order in (("0", ["garch", "polar", "beads", "rot"]), ("7", ["rot", "beads", "polar", "garch", "polar"])):
        env = dict(os.environ, PYTHONHASHSEED=seed)
        out = subprocess.run([os.sys.executable, str(script)] + order, env=env, check=True, capture_output=True, text=True)
        runs.append(dict(line.split() for line in out.stdout.strip().splitlines()))
    assert runs[0] == runs[1] and len(runs[0]) == 4


# -- tracer details ----------------------------------------------------------------------------------
def test_grad_composes_inside_traced_functions():
    u = lambda x: jnp.sum(x ** 2) * jnp.sin(x[0])
    g = pychastic.grad(u)(np.array([0.3, -0.7]))
    want = np.array([2 * 0.3 * np.sin(0.3) + (0.09 + 0.49) * np.cos(0.3), 2 * -0.7 * np.sin(0.3)])
    assert np.allclose(g, want)
    h = pychastic.hessian(u)(np.array([0.3, -0.7]))
    assert np.allclose(h, h.T) and h.shape == (2, 2)


def test_where_and_comparisons_trace():
    x = jnp.symbols((2,))
    y = jnp.where(x[0] > x[1], x[0], jnp.maximum(x[1], 0.0))
    vals = E.evaluate([E.wrap(y)], np.array([[2.0, 1.0], [1.0, 3.0], [-5.0, -1.0]]))[0]
    assert vals.tolist() == [2.0, 3.0, 0.0]
    with pytest.raises(TypeError):
        bool(x[0] > 0)


def test_rpy_mobility_matches_oracle_in_all_branches():
    radii = np.array([3.0, 1.0, 1.0, 1.0])
    for centres in ([[-2.0, 0, 0], [2.0, 0, 0], [6.0, 0, 0], [10.0, 0, 0]],         # touching + far
                    [[0.0, 0, 0], [2.5, 0.5, 0], [3.0, 1.0, 0.2], [0.5, 0.2, 0.1]]):  # overlapping + immersed
        c = np.array(centres)
        want = orpy.muTT_numpy(c, radii)
        assert np.allclose(grpy.muTT(c, radii), want, rtol=1e-13)
        x = jnp.symbols((12,))
        traced = grpy.muTT(jnp.reshape(x, (4, 3)), radii)
        vals = E.evaluate([E.wrap(v) for v in np.asarray(traced, dtype=object).ravel()], c.ravel())
        assert np.allclose(np.array([float(v) for v in vals]).reshape(12, 12), want, rtol=1e-12)


def test_symbolic_cholesky():
    rng = np.random.default_rng(2)
    A = rng.normal(size=(4, 4))
    A = A @ A.T + 4 * np.eye(4)
    x = jnp.symbols((16,))
    L = jnp.linalg.cholesky(jnp.reshape(x, (4, 4)))
    vals = E.evaluate([E.wrap(v) for v in np.asarray(L, dtype=object).ravel()], A.ravel())
    assert np.allclose(np.array([float(v) for v in vals]).reshape(4, 4), np.linalg.cholesky(A))


# -- analytic moments ----------------------------------------------------------------------------------
def test_wiener_integral_moments_known_values():
    wim = pychastic.wiener_integral_moments
    assert wim.E([1, 1])(1) == 0 and wim.E([0, 0])(1) == 0.5
    assert wim.E2([1, 1])(1) == 0.5
    assert np.isclose(wim.E2([1, 0])(1), 1 / 3) and np.isclose(wim.E2([0, 1])(1), 1 / 3)
    assert np.isclose(wim.E2([1, 0], [0, 1])(1), 1 / 6)
    assert np.isclose(wim.E2([1, 2, 1])(1), 1 / 6) and wim.E2([1, 2, 1], [2, 1, 1])(1) == 0
    assert np.isclose(wim.E2([1], [1, 0])(2.0), 2.0)      # t^2 / 2


# -- error behaviour of solve_many before any device work ----------------------------------------------------
def test_shape_mismatch_asserts_like_the_reference():
    """reference sde_solver.py:151-152: AssertionError when a(x0) / b(x0) do not match x0."""
    prob = workloads.gbm()
    prob.a = lambda x: jnp.zeros(3)
    with pytest.raises(AssertionError):
        pychastic.sde_solver.SDESolver().solve_many(prob, 2)


def test_multiple_initial_conditions_reject_n_trajectories():
    prob = workloads.gbm()
    prob.x0 = jnp.tile(prob.x0, (4, 1))
    with pytest.raises(ValueError):
        pychastic.sde_solver.SDESolver().solve_many(prob, n_trajectories=1)


def test_torch_tensors_are_accepted_as_initial_conditions():
    p = pychastic.sde_problem.SDEProblem(lambda x: -x, lambda x: jnp.eye(2), torch.tensor([1.0, 2.0]), 1.0)
    assert p.x0.tolist() == [1.0, 2.0] and p.dimension == 2


def test_lax_cond_traces_both_branches():
    from stochastic_b200 import lax
    x = jnp.symbols((1,))
    y = lax.cond(x[0] > 0.0, lambda v: v * 2.0, lambda v: -v, x[0])
    vals = E.evaluate([E.wrap(y)], np.array([[3.0], [-4.0]]))[0]
    assert vals.tolist() == [6.0, 4.0]
    assert lax.cond(True, lambda: 1.0, lambda: 2.0) == 1.0


def test_bead_ring_problem_host_side():
    prob = workloads.dna_loop(n_beads=40)
    assert prob.dimension == prob.noise_terms == 120 and prob.x0.dtype == np.float32
    par = prob.bead_ring
    assert np.isclose(par["d0"], 0.025) and np.isclose(par["ks"], 1.0 / ((0.037 / 1111.) * 0.025))
    assert np.isclose(np.linalg.norm(prob.x0[:3] - prob.x0[3:6]), 2 * (1 / (2 * np.pi)) * np.sin(np.pi / 40))
    small = workloads.dna_loop(n_beads=4, beam_length=0.1)
    a = small.a(np.asarray(small.x0, dtype=np.float64))
    b = small.b(np.asarray(small.x0, dtype=np.float64))
    assert a.shape == (12,) and b.shape == (12, 12) and np.allclose(b, np.tril(b))
    with pytest.raises(ValueError):
        workloads.BeadRingProblem(n_beads=4, x0=np.zeros(5))


def test_snapshot_tile_and_block_selection():
    from stochastic_b200 import runtime
    assert runtime.pick_snap_tile(1, 1, 128, True, 0, 512) == 16
    assert runtime.pick_snap_tile(1, 1, 128, True, 0, 1) == 1 and runtime.pick_snap_tile(1, 1, 128, True, 0, 5) == 4
    assert runtime.pick_snap_tile(2, 2, 256, True, 46 * 256 * 8, 128) == 1      # no room next to the normals staging
    assert runtime.snap_tile_bytes(1, 1, 128, True, 16) == 4 * 3 * 16 * 33 * 8


def test_host_prng_matches_oracle():
    from stochastic_b200 import prng
    from oracle import jax_prng
    assert prng.threefry2x32((0x13198A2E, 0x03707344), 0x243F6A88, 0x85A308D3) == (0xC4923A9C, 0x483DF7A0)
    assert prng.fold_in(prng.PRNGKey(0), 1) == (928981903, 3453687069)
    for seed, n in [(0, 5), (7, 1), (123456789012, 4)]:
        for layout in ("original", "partitionable"):
            for x64 in (False, True):
                assert prng.split(prng.PRNGKey(seed, x64), n, layout) == [
                    tuple(int(v) for v in k) for k in jax_prng.split(jax_prng.PRNGKey(seed, x64), n, layout)]
    from stochastic_b200.sde_solver import prng_key
    assert prng_key(5) == (0, 5) and prng_key((3, 4)) == (3, 4)
    # JAX's integer seeds are int32 unless jax_enable_x64 (jax/_src/prng.py::random_seed): the high key word is 0 and the
    # seed wraps in 32-bit mode; in 64-bit mode both words of the two's-complement seed are used
    assert prng_key(-1) == (0, 0xFFFFFFFF) and prng_key(-1, x64=True) == (0xFFFFFFFF, 0xFFFFFFFF)
    assert prng_key((1 << 33) + 5) == (0, 5) and prng_key((1 << 33) + 5, x64=True) == (2, 5)
    assert jax_prng.PRNGKey(-1).tolist() == [0, 0xFFFFFFFF] and jax_prng.PRNGKey(-1, x64=True).tolist() == [0xFFFFFFFF] * 2
    pychastic.config.update("enable_x64", True)
    try:
        assert prng_key(1 << 33) == (2, 0)
    finally:
        pychastic.config.update("enable_x64", False)


def test_warp_specialised_source_selection_and_register_budget():
    prob = workloads.polar_walk(4.0, (2.0, 0.0), f64=True)
    ks = build_source(prob.a, prob.b, 2, "wagner_platen", True, "original")
    assert "sde_kernel_ws.cuh" in ks.source and ks.block == 512 and ks.dyn_smem_bytes == 2 * 46 * 128 * 8
    assert ks.traj_per_block == 128 and ks.noise_terms == 2 and ks.n_obs == 0
    ks = build_source(prob.a, prob.b, 2, "wagner_platen", True, "original", warp_specialize=False)
    assert "sde_kernel_ws.cuh" not in ks.source and ks.block == 256 and ks.traj_per_block == 256
    ks = build_source(prob.a, prob.b, 2, "milstein", True, "original")
    assert "sde_kernel_ws.cuh" not in ks.source     # only Wagner-Platen with m > 1 stages its normals
    assert "#define SDE_MIN_BLOCKS 6" in ks.source   # small step function: 80 registers, six CTAs per SM
    prob3, _ = P.coupled_3x3()
    src = build_source(prob3.a, prob3.b, 3, "wagner_platen", True, "original")[0]
    assert "sde_kernel_ws.cuh" not in src            # ring of 2 x 69 x 128 doubles does not fit twice per SM
    with pytest.raises(ValueError):                    # 128*120 + 384*48 registers > 512*64: setmaxnreg would hang
        build_source(prob.a, prob.b, 2, "wagner_platen", True, "original",
                     ws_tuning={"SDE_WS_CONSUMER_REGS": 120, "SDE_WS_PRODUCER_REGS": 48})
    with pytest.raises(KeyError):
        build_source(prob.a, prob.b, 2, "wagner_platen", True, "original", ws_tuning={"NOPE": 1})


def test_host_side_stand_ins_for_post_processing_code():
    """`vmap` / `jit` / `tree_util.tree_map` and the numpy pass-throughs the reference's example scripts apply to
    result arrays (reference examples/hit.py:40-61, examples/jfm.py:134-203)."""
    import stochastic_b200 as jax

    def process_trajectory(traj):                       # shaped like examples/hit.py::process_trajectory
        hit = jnp.argmax(traj[:, 0] > 1.5)
        return {"hit_index": hit, "hit_place": traj[hit], "path_length": jnp.sum(jnp.abs(jnp.diff(traj[:, 0])))}

    rng = np.random.default_rng(0)
    trajs = np.cumsum(rng.normal(size=(6, 50, 2)), axis=1)
    out = jax.vmap(process_trajectory)(trajs)
    assert out["hit_index"].shape == (6,) and out["hit_place"].shape == (6, 2)
    for i in range(6):
        ref = process_trajectory(trajs[i])
        assert out["hit_index"][i] == ref["hit_index"] and np.array_equal(out["hit_place"][i], ref["hit_place"])
        assert np.isclose(out["path_length"][i], ref["path_length"])
    nested = jax.vmap(jax.vmap(lambda q: jnp.outer(q, q)))(np.ones((2, 5, 3)))
    assert nested.shape == (2, 5, 3, 3)
    assert np.array_equal(jax.vmap(lambda a, b: a * b, in_axes=(0, None))(np.arange(3.0), 2.0), [0.0, 2.0, 4.0])
    assert jax.jit(process_trajectory) is process_trajectory and jax.jit(static_argnums=0)(len) is len
    assert jax.tree_util.tree_map(lambda x: 2 * x, {"a": 1, "b": [2, 3]}) == {"a": 2, "b": [4, 6]}
    msd = np.array([0.0, 1.0, 3.0])
    assert np.array_equal(jnp.ediff1d(msd, to_end=float("nan"))[:2], [1.0, 2.0])
    assert jnp.cov(rng.normal(size=(3, 40))).shape == (3, 3)
    assert jnp.pad(np.ones((2, 2)), ((0, 0), (0, 2))).shape == (2, 4)


def test_tma_snapshot_tile_selection_and_emitted_defines():
    """Which snapshot write-out a kernel gets (csrc/sde_snap_tma.cuh): TMA tiles whenever the output rows are 16-byte
    multiples, with the longest row that keeps the kernel's occupancy; otherwise the transposed write-out loop."""
    from stochastic_b200 import runtime
    # fp64, d = m = 1: 128-byte rows fit the default budget, 64-byte rows the six-CTA budget of small step functions
    assert runtime.pick_snap_tma(1, 1, 4, True, 0, 512) == (16, 3)
    assert runtime.pick_snap_tma(1, 1, 4, True, 0, 512, budget=(227 * 1024) // 6 - 2048) == (8, 2)
    assert runtime.pick_snap_tma(1, 1, 4, False, 0, 512) == (32, 3)                      # fp32: 32 snapshots per 128-byte row
    assert runtime.pick_snap_tma(1, 1, 4, True, 0, 511) == (0, 0)                        # odd row length: no tensor map
    assert runtime.pick_snap_tma(1, 1, 4, False, 0, 510) == (0, 0)
    assert runtime.pick_snap_tma(1, 1, 4, True, 0, 6) == (4, 1)                          # short runs: rows no longer than the run
    assert runtime.pick_snap_tma(1, 1, 4, True, 0, 2) == (0, 0)
    assert runtime.pick_snap_tma(12, 12, 4, True, 0, 512) == (2, 0)                      # 25 boxes per warp: 16-byte rows
    # the warp-specialised kernel: the ring leaves room for 16-byte rows of its four consumer warps only
    ring = 2 * 46 * 128 * 8
    assert runtime.pick_snap_tma(2, 2, 4, True, ring, 512, budget=(228 // 2 - 2) * 1024) == (2, 0)
    assert runtime.snap_tma_bytes(2, 2, 4, True, 2) == 1024 + 4 * 5 * 32 * 16

    prob, _ = P.gbm()
    P.as_f64(prob)
    src = build_source(prob.a, prob.b, 1, "euler", True, "original", n_snapshots=512)
    assert "#define SDE_SNAP_TMA 8" in src.source and "#define SDE_SNAP_SWZ 2" in src.source
    assert src.dyn_smem_bytes == 1024 + 4 * 3 * 32 * 64
    odd = build_source(prob.a, prob.b, 1, "euler", True, "original", n_snapshots=511)
    assert "SDE_SNAP_TMA" not in odd.source and "#define SDE_SNAP_TILE 16" in odd.source
    final = build_source(prob.a, prob.b, 1, "euler", True, "original", n_snapshots=1)
    assert "SDE_SNAP_TMA" not in final.source and "SDE_SNAP_TILE" not in final.source
    old = pychastic.config.get("snapshot_tma")
    try:
        pychastic.config.update("snapshot_tma", False)
        loop = build_source(prob.a, prob.b, 1, "euler", True, "original", n_snapshots=512)
        assert "SDE_SNAP_TMA" not in loop.source and "#define SDE_SNAP_TILE 16" in loop.source
    finally:
        pychastic.config.update("snapshot_tma", old)
    polar, _ = P.polar_walk(4.0, (2.0, 0.0), 1.0)
    P.as_f64(polar)
    ws = build_source(polar.a, polar.b, 2, "wagner_platen", True, "original", n_snapshots=128)
    assert "sde_kernel_ws.cuh" in ws.source and "#define SDE_SNAP_TMA 2" in ws.source and "#define SDE_SNAP_SWZ 0" in ws.source
    assert ws.dyn_smem_bytes == ring + 1024 + 4 * 5 * 32 * 16


def test_library_routines_use_reciprocal_square_roots_in_fp64_kernels():
    """`jnp.linalg.cholesky` and `grpy.muTT` on traced input (config "library_rsqrt"): fp64 kernels take the pivots' and the
    pair distance's inverse powers from rsqrt alone, fp32 kernels and the switched-off configuration from sqrt + divisions
    (the one sqrt left in fp64 is the user's own: the spring length in U); the factor is the same function either way."""
    import re
    from stochastic_b200 import workloads
    from stochastic_b200.trace import emit
    prob64, prob32 = workloads.two_spheres(tmax=1.0, f64=True), workloads.two_spheres(tmax=1.0, f64=False)

    def step_text(prob, f64):
        src = build_source(prob.a, prob.b, 6, "euler", f64, "original").source
        return src[src.index("sde_step"):]

    def count(text, fn):
        return len(re.findall(r"\b" + fn + r"\(", text))

    s64 = step_text(prob64, True)
    n_piv = count(s64, "rsqrt")                 # (equal pivots - the three of a sphere's own block - are one expression)
    assert 4 <= n_piv <= 7 and count(s64, "sqrt") == 1
    s32 = step_text(prob32, False)
    assert count(s32, "rsqrt") == 0 and count(s32, "sqrt") >= n_piv - 1
    old = pychastic.config.get("library_rsqrt")
    try:
        pychastic.config.update("library_rsqrt", False)
        off = step_text(prob64, True)
    finally:
        pychastic.config.update("library_rsqrt", old)
    assert count(off, "rsqrt") == 0 and count(off, "sqrt") == count(s32, "sqrt")
    assert off.count(" / ") > s64.count(" / ")                   # and the pivots' divisions
    # same function: the symbolic factor evaluated on the host at a random configuration
    from stochastic_b200 import numpy as jnp_
    from stochastic_b200.trace import expr as E
    rng = np.random.default_rng(3)
    G = rng.standard_normal((5, 7))
    C = G @ G.T + np.eye(5)
    for flag in (True, False):
        pychastic.config.update("library_rsqrt", flag)
        try:
            xs = [E.var(k) for k in range(25)]
            L = jnp_.linalg.cholesky(np.array(xs, dtype=object).reshape(5, 5).view(type(jnp_.asarray(xs))))
            vals = E.evaluate([E.wrap(v) for v in np.asarray(L, dtype=object).ravel()], C.ravel())
        finally:
            pychastic.config.update("library_rsqrt", old)
        assert np.allclose(np.array(vals).reshape(5, 5), np.linalg.cholesky(C), rtol=1e-13, atol=1e-15)


def test_bead_ring_kernel_selection_by_size_and_dtype():
    """Which bead-chain kernel a ring gets: the register-tiled one up to 3 N = 128 in either dtype (fp32 problems with fp32
    in / out around its fp64 arithmetic), the shared-memory one beyond or on request; small rings are traced unless forced."""
    from stochastic_b200 import workloads
    solver = pychastic.sde_solver.SDESolver(scheme="euler", dt=1e-7)
    for f64 in (True, False):
        src = solver._bead_plan(workloads.dna_loop(n_beads=40, f64=f64), f64, "original").ks.source
        assert "#define SDE_BEAD_REGTILE 1" in src and "#define SDE_F64 1" in src and "#define SDE_BLOCK 256" in src
        assert f"#define SDE_BEAD_IO_F32 {0 if f64 else 1}" in src
    big = solver._bead_plan(workloads.dna_loop(n_beads=48, f64=False), False, "original").ks.source      # 3 N = 144
    assert "#define SDE_BEAD_REGTILE 0" in big and "#define SDE_F64 0" in big and "#define SDE_BEAD_IO_F32 0" in big
    solver.bead_regtile = False
    off = solver._bead_plan(workloads.dna_loop(n_beads=40, f64=False), False, "original").ks.source
    assert "#define SDE_BEAD_REGTILE 0" in off and "#define SDE_F64 0" in off
    solver.bead_regtile = True
    with pytest.raises(ValueError):
        solver._bead_plan(workloads.dna_loop(n_beads=48, f64=True), True, "original")
    assert not solver._uses_bead_kernel(workloads.dna_loop(n_beads=5, beam_length=0.125), False)
    assert solver._uses_bead_kernel(workloads.dna_loop(n_beads=5, beam_length=0.125), True)
    assert solver._uses_bead_kernel(workloads.dna_loop(n_beads=40), False)
