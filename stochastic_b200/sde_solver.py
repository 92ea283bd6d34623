"""`SDESolver`: batched Euler / Milstein / Wagner-Platen integration on B200.

Drop-in for the reference's `pychastic.sde_solver.SDESolver`
(reference `pychastic/sde_solver.py:49-359`): same constructor, `solve_many` / `solve`
signatures, result dict (`time_values`, `solution_values`, `wiener_values`; snapshots exclude
t=0) and error behaviour.  What changes is everything underneath:

* the vmapped 3-level `lax.scan` nest (`sde_solver.py:228-300`) is ONE launch of a
  thread-per-trajectory kernel (csrc/sde_kernel.cuh) whose step loop, state and random
  stream live in registers;
* `jax.random` + `get_wiener_integrals` (`vectorized_I_generation.py`) are generated in-kernel,
  bit-compatible with the reference's key fan-out (csrc/sde_core.cuh);
* the Taylor-Ito coefficient closures (`sde_solver.py:157-184`) are differentiated symbolically
  at trace time and emitted as one CUDA device function, NVRTC-compiled for sm_100a.

Host-side integer/float arithmetic follows the reference exactly (`int(tmax/dt)`, the chunk
round-up that overshoots tmax, dt**1.5 in Python double).
"""
import ctypes
import os
from collections import OrderedDict
from functools import wraps
from typing import NamedTuple

import numpy as np

from . import config, jnp, prng, runtime
from .autodiff import hessian, jacobian
from .device_array import DeviceArray
from .sde_problem import SDEProblem
from .trace import emit
from .trace import expr as E
from .trace.taylor import SCHEMES, TracedProblem, step_expressions

SCHEME_ID = {"euler": 0, "milstein": 1, "wagner_platen": 2}


class KernelSource(NamedTuple):
    """What `build_source` hands to the runtime: the translation unit and its launch geometry."""
    source: str
    noise_terms: int
    n_obs: int
    block: int              # threads per CTA (SDE_BLOCK)
    dyn_smem_bytes: int
    traj_per_block: int     # trajectories per CTA: block, or 128 for the warp-specialised kernel


# ------------------------------------------------------------------------------------------
# Taylor-Ito operators with the reference's calling convention (imported by its tests:
# reference tests/test_coefficient_function_logic_1d.py:15-16)
# ------------------------------------------------------------------------------------------
def contract_all(a, b):
    return jnp.tensordot(a, b, axes=len(np.shape(b)))


def tensordot1(a, b):
    return jnp.tensordot(a, b, axes=1)


def tensordot2(a, b):
    return jnp.tensordot(a, b, axes=2)


def L_t_operator(f, problem):
    """L^0 f = df.a + 1/2 d2f:(b b^T); result indexed [spatial, time(=1), ...]  (reference :29-38)."""
    @wraps(f)
    def wrapped(x):
        b_val = problem.b(x)
        val = tensordot1(jacobian(f)(x), problem.a(x)) + 0.5 * tensordot2(
            hessian(f)(x), tensordot1(b_val, jnp.transpose(b_val)))
        return val[:, None, ...]
    return wrapped


def L_w_operator(f, problem):
    """L^j f = df.b; result indexed [spatial, noise, ...]  (reference :40-46)."""
    @wraps(f)
    def wrapped(x):
        val = tensordot1(jacobian(f)(x), problem.b(x))[:, None, ...]
        return jnp.swapaxes(val, 1, -1)[:, ..., 0]
    return wrapped


# ------------------------------------------------------------------------------------------
# host arithmetic shared with the tests
# ------------------------------------------------------------------------------------------
def chunk_arithmetic(tmax, dt, chunk_size, chunks_per_randomization):
    """(chunk_size, number_of_chunks, chunks_per_randomization, n_fragments), reference :217-221."""
    steps_needed = int(tmax / dt)
    tmp = chunks_per_randomization or 1
    chunk_size = chunk_size or steps_needed
    number_of_chunks = ((steps_needed // (chunk_size * tmp)) + (1 if steps_needed % (chunk_size * tmp) else 0)) * tmp
    chunks_per_randomization = chunks_per_randomization or number_of_chunks
    return chunk_size, number_of_chunks, chunks_per_randomization, number_of_chunks // chunks_per_randomization


def prng_key(seed, x64=None):
    """Raw key data of jax.random.PRNGKey(seed) (`prng.PRNGKey`: the high word is 0 unless the run is in 64-bit
    mode, as in JAX); a (hi, lo) pair of uint32 words is passed through unchanged, so that a key derived on the
    host (e.g. `prng.fold_in(PRNGKey(seed), rank)`) can be used as `seed`."""
    if isinstance(seed, (tuple, list, np.ndarray)) and len(seed) == 2:
        return int(seed[0]) & 0xFFFFFFFF, int(seed[1]) & 0xFFFFFFFF
    return prng.PRNGKey(seed, x64)


def _c_function(name, signature, outputs, var_names, f64):
    body = emit.emit_body(outputs, var_names, f64=f64)
    return f"__device__ __forceinline__ void {name}({signature}) {{\n{body}\n}}\n"


def _emit_post(post, d, f64):
    x = jnp.symbols((d,))
    y = np.asarray(jnp._as_obj(post(x)), dtype=object)
    if y.shape != (d,):
        raise ValueError(f"step_post_processing returned shape {y.shape}, expected {(d,)}")
    names = {i: f"x[{i}]" for i in range(d)}
    return _c_function("sde_post", "real (&x)[SDE_D]", [(f"x[{i}]", E.wrap(v)) for i, v in enumerate(y)], names, f64), d


def observe_arity(observe):
    """3: observables(x, w, t); 4: observables(x, w, t, x_ref) with x_ref = the state at the first snapshot."""
    import inspect
    try:
        params = [p for p in inspect.signature(observe).parameters.values()
                  if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD) and p.default is p.empty]
    except (TypeError, ValueError):
        return 3
    if len(params) not in (3, 4):
        raise ValueError("observables must take (x, w, t) or (x, w, t, x_ref)")
    return len(params)


def _emit_observe(observe, d, m, f64):
    x = jnp.symbols((d,))
    w = jnp.symbols((m,), start=d)
    t = E.var(d + m)
    uses_ref = observe_arity(observe) == 4
    if uses_ref:
        xr = jnp.symbols((d,), start=d + m + 1)
        o = observe(x, w, t, xr)
    else:
        o = observe(x, w, t)
    o = np.atleast_1d(np.asarray(jnp._as_obj(o), dtype=object)).ravel()
    names = {i: f"x[{i}]" for i in range(d)}
    names.update({d + j: f"w[{j}]" for j in range(m)})
    names[d + m] = "t"
    names.update({d + m + 1 + i: f"xr[{i}]" for i in range(d)})
    src = _c_function("sde_observe", "const real* x, const real* w, const real t, const real* xr, "
                      "double (&obs)[SDE_NOBS]", [(f"obs[{k}]", E.wrap(v)) for k, v in enumerate(o)], names, f64)
    return src, o.size, uses_ref


def _emit_event(event, d, m, f64):
    x = jnp.symbols((d,))
    w = jnp.symbols((m,), start=d)
    t = E.var(d + m)
    g = np.asarray(jnp._as_obj(event(x, w, t)), dtype=object).ravel()
    if g.size != 1:
        raise ValueError(f"first_passage must return one scalar, got {g.size} values")
    names = {i: f"x[{i}]" for i in range(d)}
    names.update({d + j: f"w[{j}]" for j in range(m)})
    names[d + m] = "t"
    body = emit.emit_body([("ev", E.wrap(g[0]))], names, f64=f64)
    return ("__device__ __forceinline__ real sde_event(const real* x, const real* w, const real t) {\n"
            f"    real ev;\n{body}\n    return ev;\n}}\n")


def _emit_hooks(parts, d, m, f64, post, observe, event, observe_at):
    """Optional device functions of a stepping kernel (sde_post / sde_observe / sde_event) and their macros."""
    n_obs = 0
    post_src = obs_src = ev_src = ""
    if post is not None:
        post_src, _ = _emit_post(post, d, f64)
        parts.append("#define SDE_HAS_POST 1")
    if observe is not None:
        obs_src, n_obs, uses_ref = _emit_observe(observe, d, m, f64)
        parts.append(f"#define SDE_NOBS {n_obs}")
        if observe_at == "snapshots":
            parts.append("#define SDE_OBS_SNAP 1")
        elif observe_at != "final":
            raise ValueError(f"observe_at must be 'final' or 'snapshots', not {observe_at!r}")
        if uses_ref:
            parts.append("#define SDE_OBS_REF 1")
    if event is not None:
        ev_src = _emit_event(event, d, m, f64)
        parts.append("#define SDE_HAS_EVENT 1")
    return n_obs, [post_src, obs_src, ev_src]


def build_source(problem_a, problem_b, d, scheme, f64, layout, post=None, observe=None, p=10, block=None,
                 min_blocks=None, lockstep=None, n_snapshots=1, snap_tile=None, warp_specialize=None, ws_tuning=None,
                 event=None, observe_at="final"):
    """Trace the problem and emit the complete translation unit of its stepping kernel.

    Returns a `KernelSource`."""
    E.trim_table()
    E.TRACE_F64 = bool(f64)
    tp = TracedProblem(problem_a, problem_b, d)
    m = tp.m
    outs, sv = step_expressions(tp, scheme)
    stage = runtime.stage_count(m, p, scheme)
    elem = 8 if f64 else 4
    if warp_specialize is None:
        warp_specialize = config.get("warp_specialize")
    ws = bool(warp_specialize) and scheme == "wagner_platen" and m > 1 and 2 * stage * 128 * elem <= 110 * 1024
    step_src = lambda: _c_function("sde_step", "real (&x)[SDE_D], const sdeb::Integrals& I, const real dt",
                                   [(f"x[{i}]", e) for i, e in enumerate(outs)], sv.names, f64)
    if ws:
        # warp-specialised kernel (csrc/sde_kernel_ws.cuh): 128 consumer + 384 producer threads, two-stage ring
        parts = [
            "// generated by stochastic_b200.sde_solver.build_source (warp-specialised)",
            f"#define SDE_D {d}", f"#define SDE_M {m}", f"#define SDE_SCHEME {SCHEME_ID[scheme]}",
            f"#define SDE_P {int(p)}", f"#define SDE_F64 {1 if f64 else 0}",
            f"#define SDE_LAYOUT {1 if layout == 'partitionable' else 0}",
            "#define SDE_STAGE_STRIDE 128",
        ]
        # measured on B200 (profiles/README.md): 3 producer warpgroups per consumer warpgroup, one normal at a time
        knobs = {"SDE_WS_PGROUPS": 3, "SDE_WS_CONSUMER_REGS": 136, "SDE_WS_PRODUCER_REGS": 40, "SDE_WS_GROUP": 1,
                 "SDE_WS_CTAIL": 0, "SDE_LOGTAB_IN_SMEM": 1, "SDE_WS_SPLIT_NORMAL": 0}
        for knob, val in (ws_tuning or {}).items():
            if knob not in knobs:
                raise KeyError(f"unknown warp-specialisation knob {knob!r}")
            knobs[knob] = int(val)
        ws_block = 128 * (1 + knobs["SDE_WS_PGROUPS"])
        launch_regs = (65536 // (2 * ws_block)) // 8 * 8            # __launch_bounds__(ws_block, 2)
        need = 128 * knobs["SDE_WS_CONSUMER_REGS"] + 128 * knobs["SDE_WS_PGROUPS"] * knobs["SDE_WS_PRODUCER_REGS"]
        if need > launch_regs * ws_block or knobs["SDE_WS_CONSUMER_REGS"] % 8 or knobs["SDE_WS_PRODUCER_REGS"] % 8:
            raise ValueError(f"warp-specialised register split {knobs} exceeds the CTA's {launch_regs * ws_block} "
                             "registers (setmaxnreg.inc would never be granted)")
        parts.append(f"#define SDE_BLOCK {ws_block}")
        for knob, val in knobs.items():
            parts.append(f"#define {knob} {val}")
        if knobs["SDE_WS_SPLIT_NORMAL"]:
            parts.append("#define SDE_TAIL_ATTR __forceinline__")        # (same ptxas issue as below)
        if knobs["SDE_WS_CTAIL"]:
            if not 0 < knobs["SDE_WS_CTAIL"] <= 3 * m or knobs["SDE_WS_GROUP"] != 1:
                raise ValueError("SDE_WS_CTAIL must be in [0, 3m] and needs SDE_WS_GROUP == 1")
            # ptxas 12.9 crashes on a call to the out-of-line erf_inv tail from the register-grown consumer branch
            parts.append("#define SDE_TAIL_ATTR __forceinline__")
        ring_bytes = 2 * stage * 128 * elem
        # stored trajectories: snapshot tiles of the four consumer warps behind the ring, as large as two CTAs per SM
        # leave room for (228 KB per SM, 1 KB reserved per CTA, 1 KB of static shared memory for the log table)
        tma_tile = tma_swz = 0
        if isinstance(snap_tile, tuple):
            _, tma_tile, tma_swz = snap_tile
        elif snap_tile is None and config.get("snapshot_tma"):
            tma_tile, tma_swz = runtime.pick_snap_tma(d, m, 4, f64, ring_bytes, n_snapshots, budget=(228 // 2 - 2) * 1024)
        if tma_tile:
            parts.append(f"#define SDE_SNAP_TMA {int(tma_tile)}")
            parts.append(f"#define SDE_SNAP_SWZ {int(tma_swz)}")
            ring_bytes += runtime.snap_tma_bytes(d, m, 4, f64, tma_tile)
        n_obs, hooks = _emit_hooks(parts, d, m, f64, post, observe, event, observe_at)
        parts.append('#include "sde_core.cuh"')
        parts.append(f'static_assert(SDE_STAGE_COUNT == {stage}, "host/device staging size mismatch");')
        parts.append(step_src())
        parts += hooks + ['#include "sde_kernel_ws.cuh"']
        return KernelSource("\n".join(parts) + "\n", m, n_obs, ws_block,
                            ring_bytes, 128)
    user_block = block
    block = runtime.pick_block(m, p, scheme, f64, block)
    if user_block is None and min_blocks is None and stage and scheme == "wagner_platen":
        # Wagner-Platen outside the warp-specialised kernel (three or more noise terms): step functions of many thousand
        # statements (rotational Brownian motion: 16 000) live in local memory at 128 registers - 128-thread CTAs at up to
        # 255 registers run them 40 % faster (profiles/README.md)
        if sum(1 for node in E.topo_order(outs) if node.op not in ("const", "var")) > 4000:
            block, min_blocks = 128, 2
    if min_blocks is None and stage and block == 256:
        # measured on B200 (profiles/README.md): two 256-thread CTAs per SM at <= 128 registers beat one CTA at 244
        min_blocks = 2
    if min_blocks is None and stage and block == 128 and scheme == "euler":
        min_blocks = 3                # (see runtime.pick_block)
    if min_blocks is None and not stage and block == 128:
        # small step functions: cap the kernel at 80 registers so that six CTAs are resident per SM - the normal
        # generator is latency-bound and gains more from extra warps than from registers (two spheres +20 %, polar
        # Euler +55 %, GARCH Milstein +10 %; the 1 640-node four-bead step function loses 20 % and is left alone)
        n_nodes = sum(1 for node in E.topo_order(outs) if node.op not in ("const", "var"))
        if n_nodes <= 800:
            min_blocks = 6
    parts = [
        "// generated by stochastic_b200.sde_solver.build_source",
        f"#define SDE_D {d}", f"#define SDE_M {m}", f"#define SDE_SCHEME {SCHEME_ID[scheme]}",
        f"#define SDE_P {int(p)}", f"#define SDE_F64 {1 if f64 else 0}",
        f"#define SDE_LAYOUT {1 if layout == 'partitionable' else 0}", f"#define SDE_BLOCK {int(block)}",
    ]
    stage_bytes = stage * block * elem
    # snapshot write-out: tiles through the TMA unit when the output rows can be described by a tensor map
    # (snap_tile = None or ("tma", tile, swizzle)), else a transposed write-out of `snap_tile` staged snapshots, else
    # direct stores
    tma_tile = tma_swz = 0
    if snap_tile is None and config.get("snapshot_tma"):
        # kernels capped at six CTAs per SM (small step functions) keep that occupancy: measured on B200, C1 with every
        # step stored runs at 3.67 TB/s with 64-byte rows and six CTAs against 3.42 TB/s with 128-byte rows and four
        budget = (227 * 1024) // 6 - 2048 if min_blocks == 6 else 96 * 1024
        tma_tile, tma_swz = runtime.pick_snap_tma(d, m, block // 32, f64, stage_bytes, n_snapshots, budget=budget)
    elif isinstance(snap_tile, tuple):
        _, tma_tile, tma_swz = snap_tile
    if tma_tile:
        snap_tile = 1
        parts.append(f"#define SDE_SNAP_TMA {int(tma_tile)}")
        parts.append(f"#define SDE_SNAP_SWZ {int(tma_swz)}")
    elif snap_tile is None:
        snap_tile = runtime.pick_snap_tile(d, m, block, f64, stage_bytes, n_snapshots)
    if snap_tile > 1:
        parts.append(f"#define SDE_SNAP_TILE {int(snap_tile)}")
    if min_blocks:
        parts.append(f"#define SDE_MIN_BLOCKS {int(min_blocks)}")
    if lockstep is not None:
        parts.append(f"#define SDE_LOCKSTEP {1 if lockstep else 0}")
    n_obs, hooks = _emit_hooks(parts, d, m, f64, post, observe, event, observe_at)
    parts.append('#include "sde_core.cuh"')
    parts.append(f'static_assert(SDE_STAGE_COUNT == {stage}, "host/device staging size mismatch");')
    parts.append(step_src())
    parts += hooks
    parts.append('#include "sde_kernel.cuh"')
    return KernelSource("\n".join(parts) + "\n", m, n_obs, block,
                        stage_bytes + runtime.snap_tile_bytes(d, m, block, f64, snap_tile) +
                        runtime.snap_tma_bytes(d, m, block // 32, f64, tma_tile), block)


def bead_regtile_block(n_beads):
    """Threads per CTA of the register-tiled bead kernel: one chain warp and NT / 2 bulk warps (NT = tile rows of 8)."""
    nt = -(-3 * n_beads // 8)
    return 32 * (1 + max(1, nt // 2))


def bead_regtile_smem_bytes(n_beads, twist=False):
    """Dynamic shared memory of the register-tiled bead kernel; mirrors RT_SMEM_DOUBLES of csrc/sde_bead_kernel.cuh."""
    npad = -(-3 * n_beads // 8) * 8
    nt = npad // 8
    npair = n_beads * (n_beads - 1) // 2
    dch = max(3, min(6, bead_regtile_block(n_beads) // n_beads))          # drift chunks per bead (RT_DCH)
    twsc = (n_beads // 2 - 1) * n_beads * 6 if twist else 0               # writhe gradients of the second segments (RT_TWSC)
    return 8 * ((8 + dch) * npad + (4 + 6 * npair) + npair + (npair & 1) + 2 * nt * 64 + 128 + twsc)


def bead_ring_source(n_beads, f64, layout, block, twist=False, regtile=False):
    """Translation unit of the CTA-per-trajectory bead-chain kernel (csrc/sde_bead_kernel.cuh)."""
    return "\n".join([
        "// generated by stochastic_b200.sde_solver (bead-chain functor)",
        f"#define SDE_NB {n_beads}", f"#define SDE_F64 {1 if (f64 or regtile) else 0}",
        f"#define SDE_BEAD_IO_F32 {1 if (regtile and not f64) else 0}",   # fp32 problem on the fp64 register-tiled kernel
        f"#define SDE_LAYOUT {1 if layout == 'partitionable' else 0}", f"#define SDE_BLOCK {block}",
        f"#define SDE_BEAD_DMMA {1 if (f64 or regtile) else 0}",       # fp64: trailing update on the FP64 tensor cores
        f"#define SDE_BEAD_TWIST {1 if twist else 0}",    # writhe-dependent twist energy (reference :57)
        f"#define SDE_BEAD_REGTILE {1 if regtile else 0}",  # mobility / Cholesky factor in tensor-core accumulator tiles
        f"#define SDE_BEAD_CHAIN_PER_THREAD {int(os.environ.get('SDEB200_BEAD_CHAIN_PER_THREAD', '0'))}",
        f"#define SDE_BEAD_TIMING {int(os.environ.get('SDEB200_BEAD_TIMING', '0'))}",   # diagnostics: per-phase cycle counts
        *([f"#define SDE_BEAD_EXPERIMENT_SKIP {int(os.environ['SDEB200_BEAD_SKIP'])}"] if os.environ.get("SDEB200_BEAD_SKIP") else []),
        *([f"#define SDE_BEAD_UPD_GROUP {int(os.environ['SDEB200_BEAD_UPD_GROUP'])}"] if os.environ.get("SDEB200_BEAD_UPD_GROUP") else []),
        *([f"#define SDE_BEAD_FACTOR_DMMA {int(os.environ['SDEB200_BEAD_FACTOR_DMMA'])}"] if os.environ.get("SDEB200_BEAD_FACTOR_DMMA") else []),
        *([f"#define SDE_BEAD_FENCES {int(os.environ['SDEB200_BEAD_FENCES'])}"] if os.environ.get("SDEB200_BEAD_FENCES") else []),
        '#include "sde_bead_kernel.cuh"', ""])


class _Plan:
    """One compiled-kernel choice of `SDESolver`: the emitted source and its launch geometry, plus the loaded
    program per device.  Cached per (callables, scheme, dtype, knobs), so that a repeated `solve_many` on the same
    problem neither traces, differentiates, emits nor hashes anything again."""
    __slots__ = ("ks", "entry", "programs", "keepalive")

    def __init__(self, ks, entry, keepalive):
        self.ks, self.entry, self.programs, self.keepalive = ks, entry, {}, keepalive

    def program(self):
        """The kernel loaded on the current device (raises without a CUDA device)."""
        if os.environ.get("SDEB200_PRECOMPILE_ONLY"):
            return runtime.load_program(self.ks.source, self.entry)
        torch = runtime.require_cuda()
        dev = torch.cuda.current_device()
        prog = self.programs.get(dev)
        if prog is None:
            prog = self.programs[dev] = runtime.load_program(self.ks.source, self.entry)
        return prog


_PLANS = OrderedDict()
_PLAN_LIMIT = 256
PLAN_STATS = {"hits": 0, "misses": 0}


def _cached_plan(key, keepalive, make):
    plan = _PLANS.get(key)
    if plan is not None:
        _PLANS.move_to_end(key)
        PLAN_STATS["hits"] += 1
        return plan
    PLAN_STATS["misses"] += 1
    ks, entry = make()
    # the key holds ids of the user's callables: keep them alive as long as the entry, so an id is never reused
    plan = _PLANS[key] = _Plan(ks, entry, keepalive)
    while len(_PLANS) > _PLAN_LIMIT:
        _PLANS.popitem(last=False)
    return plan


class SDESolver:
    """
    Parameters
    ----------
    scheme : {'euler', 'milstein', 'wagner_platen'}, default 'euler'
    dt : float, step size of the fixed-step integration.

    The remaining keyword arguments are accepted for compatibility with the reference, which
    stores them and never reads them (reference `sde_solver.py:63-80`).
    """

    def __init__(self, scheme="euler", dt=0.01, dt_adapting_factor=10, min_dt=None, max_dt=None,
                 error_terms=1, target_mse_density=1e-2, adaptive=False):
        self.scheme = scheme
        self.dt = dt
        self.min_dt = min_dt or self.dt / dt_adapting_factor
        self.max_dt = max_dt or self.dt * dt_adapting_factor
        self.error_terms = error_terms
        self.target_mse_density = target_mse_density
        self.adaptive = adaptive
        # kernel tuning knobs (None: chosen per kernel, see runtime.pick_block / csrc/sde_kernel.cuh)
        self.block_size = None
        self.min_blocks_per_sm = None
        self.lockstep = None
        self.snap_tile = None
        self.warp_specialize = None    # None: config "warp_specialize"; producer/consumer kernel (sde_kernel_ws.cuh)
        self.ws_tuning = None          # e.g. {"SDE_WS_PGROUPS": 2, "SDE_WS_CONSUMER_REGS": 160, ...}
        self.use_bead_kernel = False   # force the CTA-per-trajectory kernel for BeadRingProblem of any size
        self.bead_regtile = None       # register-tiled bead kernel: None = whenever it applies (3 N <= 128), see _bead_plan

    # -- kernel selection (shared by compile and solve_many) -----------------------------------
    def _knobs(self):
        return (self.block_size, self.min_blocks_per_sm, self.lockstep, self.snap_tile, self.warp_specialize,
                tuple(sorted((self.ws_tuning or {}).items())), self.use_bead_kernel, self.bead_regtile,
                config.get("warp_specialize"), config.get("reciprocal_division"), config.get("snapshot_tma"),
                config.get("library_rsqrt"))

    @staticmethod
    def _uses_bead_kernel(problem, force):
        return getattr(problem, "bead_ring", None) is not None and (problem.dimension > 24 or force)

    def _bead_plan(self, problem, f64, layout):
        par = problem.bead_ring
        nb = par["n_beads"]
        block = max(128, -(-3 * nb // 32) * 32)
        if self.block_size:
            if self.block_size % 32 or not block <= self.block_size <= 1024:
                raise ValueError(f"bead kernel: block_size must be a multiple of 32 in [{block}, 1024]")
            block = int(self.block_size)
        twist = bool(par.get("twist"))
        # Register-tiled variant (mobility / Cholesky factor in tensor-core accumulator fragments, fraction-free pivot
        # chain, row solves as tensor-core products): the default wherever it applies.  Measured on B200
        # (profiles/README.md, 8 880 x 100 steps): 1.07e7 against 7.3e6 trajectory-steps/s without the twist term,
        # 6.4e6 against 5.3e6 with it.
        # fp32 problems run on it too: fp32 initial condition / snapshots / random draws, fp64 in between (SDE_BEAD_IO_F32)
        regtile = (3 * nb <= 128) if self.bead_regtile is None else bool(self.bead_regtile)
        if regtile:
            if not 3 * nb <= 128:
                raise ValueError("bead_regtile needs 3 N <= 128")
            if self.block_size and self.block_size != bead_regtile_block(nb):
                raise ValueError(f"bead_regtile: the block size is fixed by the ring size ({bead_regtile_block(nb)} threads)")
            block = bead_regtile_block(nb)
        key = ("bead", nb, f64, layout, block, twist, regtile)
        return _cached_plan(key, (), lambda: (KernelSource(bead_ring_source(nb, f64, layout, block, twist, regtile),
                                                           3 * nb, 0, block, 0, 2 if regtile else 1), "sde_bead_kernel"))

    def _plan(self, problem, d, f64, layout, post, observe, event, n_snapshots, observe_at, first_sight=None):
        """The (cached) stepping kernel for this problem and these options; `first_sight()` runs before the problem
        is traced for the first time."""
        # all that pick_snap_tile / pick_snap_tma read of the snapshot count
        snap_class = (min(int(n_snapshots), 32), (int(n_snapshots) * (8 if f64 else 4)) % 16 == 0)
        key = ("step", id(problem.a), id(problem.b), d, self.scheme, f64, layout, id(post) if post else None,
               id(observe) if observe else None, id(event) if event else None, observe_at, snap_class, self._knobs())
        return _cached_plan(key, (problem.a, problem.b, post, observe, event), lambda: (
            first_sight and first_sight(),
            build_source(problem.a, problem.b, d, self.scheme, f64, layout, post=post, observe=observe,
                         block=self.block_size, min_blocks=self.min_blocks_per_sm, lockstep=self.lockstep,
                         n_snapshots=n_snapshots, snap_tile=self.snap_tile, warp_specialize=self.warp_specialize,
                         ws_tuning=self.ws_tuning, event=event, observe_at=observe_at), "sde_solve_kernel")[1:])

    def compile(self, problem, step_post_processing=None, observables=None, prng_layout=None, chunk_size=1,
                chunks_per_randomization=None, max_snapshots=None, first_passage=None, store_trajectories=True,
                observe_at="final"):
        """Trace `problem`, emit the kernel that `solve_many` would launch with the same keyword arguments and
        NVRTC-compile it into the cubin cache (no GPU needed).  Returns the cache key."""
        if self.scheme not in SCHEMES:
            raise ValueError(f"unknown scheme {self.scheme!r}; expected one of {SCHEMES}")
        x0 = np.asarray(jnp._host(problem.x0))
        f64 = x0.dtype == np.float64
        d = x0.shape[-1]
        layout = prng_layout or config.get("prng_layout")
        if self._uses_bead_kernel(problem, self.use_bead_kernel):
            plan = self._bead_plan(problem, f64, layout)
        else:
            _, n_chunks, _, _ = chunk_arithmetic(problem.tmax, self.dt, chunk_size, chunks_per_randomization)
            n_snap = n_chunks if max_snapshots is None else min(n_chunks, int(max_snapshots))
            plan = self._plan(problem, d, f64, layout, step_post_processing, observables, first_passage,
                              n_snap if store_trajectories else 1, observe_at)
        return runtime.compile_source(plan.ks.source)[0]

    # -- the hot path ------------------------------------------------------------------------
    def solve_many(self, problem: SDEProblem, n_trajectories=1, step_post_processing=None, seed=0, chunk_size=1,
                   chunks_per_randomization=None, progress_bar=True, *, observables=None, first_passage=None,
                   store_trajectories=True, traj_range=None, max_snapshots=None, prng_layout=None,
                   observe_at="final"):
        """Integrate `n_trajectories` sample paths (reference `sde_solver.py:82-304`).

        Extensions (keyword-only, not in the reference):
        observables        callable (x, w, t) -> vector, or (x, w, t, x_ref) -> vector with x_ref the state at the
                           first snapshot (what the reference's mean-square-displacement post-processing subtracts,
                           `examples/jfm.py:134-171`); its per-trajectory values are reduced on the device to sums /
                           sums of squares (result key 'moments', plus 'count').
        observe_at         'final' (default): observables are evaluated once, at the last snapshot; 'moments' has
                           shape (n_obs, 2).  'snapshots': at every snapshot; shape (n_snapshots, n_obs, 2) - the
                           ensemble statistics of a published study without ever storing (n_traj, n_snap, d).
        first_passage      callable (x, w, t) -> scalar event function; for every trajectory the first step after which
                           it is positive is located by linear interpolation between the states before and after
                           that step, on the device (result keys 'hit_flag', 'hit_time', 'hit_state'; the host
                           post-processing of reference `examples/hit.py:40-58` without storing the trajectory).
        store_trajectories False: do not allocate or write the snapshot arrays (moments-only run).
        traj_range         (begin, count): integrate only this slice of the n_trajectories-wide key
                           fan-out (multi-GPU sharding; keys stay those of the full run).
        max_snapshots      stop after this many snapshots (prefix of a long run, same keys).
        prng_layout        'original' | 'partitionable' (default: config 'prng_layout').
        """
        scheme = self.scheme
        if scheme not in SCHEMES:
            raise ValueError(f"unknown scheme {scheme!r}; expected one of {SCHEMES}")
        if self._uses_bead_kernel(problem, self.use_bead_kernel):
            if step_post_processing is not None or observables is not None or first_passage is not None:
                raise NotImplementedError("the bead-chain kernel supports neither step_post_processing, observables "
                                          "nor first_passage")
            return self._solve_bead_ring(problem, n_trajectories, seed, chunk_size, chunks_per_randomization,
                                         traj_range, max_snapshots, prng_layout)
        x0 = np.asarray(jnp._host(problem.x0))
        multiple_ic = x0.ndim == 2
        if multiple_ic and n_trajectories is not None:
            raise ValueError("n_trajectories option not supproted with multiple inital conditions!")
        if not np.issubdtype(x0.dtype, np.floating):
            x0 = x0.astype(np.float32)
        f64 = x0.dtype == np.float64
        if not f64:
            x0 = x0.astype(np.float32)
        ics = x0 if multiple_ic else x0.reshape(1, -1)
        n_total = ics.shape[0] if multiple_ic else int(n_trajectories)
        d = ics.shape[1]

        layout = prng_layout or config.get("prng_layout")
        chunk, n_chunks, cpr, n_frag = chunk_arithmetic(problem.tmax, self.dt, chunk_size, chunks_per_randomization)
        n_snap = n_chunks if max_snapshots is None else min(n_chunks, int(max_snapshots))

        def shape_assertions():      # the reference's (`sde_solver.py:151-152`), once per set of callables
            a0 = np.shape(jnp._as_obj(problem.a(ics[0])))
            b0 = np.shape(jnp._as_obj(problem.b(ics[0])))
            assert ics[0].shape == a0
            assert ics[0].shape[0] == b0[0]

        plan = self._plan(problem, d, f64, layout, step_post_processing, observables, first_passage,
                          n_snap if store_trajectories else 1, observe_at, first_sight=shape_assertions)
        source, m, n_obs, block, dyn_smem, traj_per_block = plan.ks
        prog = plan.program()   # raises without a CUDA device
        torch = runtime.require_cuda()

        S = chunk * cpr
        largest_draw = S * (10 * m if (m > 1 and scheme != "euler") else max(m, 2))
        if 2 * largest_draw >= 2 ** 32:
            raise ValueError("chunk_size * chunks_per_randomization too large for the 32-bit counter layout; "
                             "set chunks_per_randomization")
        if layout == "original" and 2 * n_total > 2 ** 32:
            raise ValueError("n_trajectories too large for the original key layout (2 n <= 2^32); "
                             "use prng_layout='partitionable'")
        begin, count = (0, n_total) if traj_range is None else (int(traj_range[0]), int(traj_range[1]))
        if begin < 0 or count < 0 or begin + count > n_total:
            raise ValueError(f"traj_range {traj_range} outside [0, {n_total})")

        dev = torch.device("cuda", torch.cuda.current_device())
        tdt = torch.float64 if f64 else torch.float32
        ic_host = ics[begin:begin + count] if multiple_ic else ics
        x0_dev = torch.as_tensor(np.ascontiguousarray(ic_host), dtype=tdt).to(dev, non_blocking=True)
        out_t = out_x = out_w = None
        if store_trajectories:
            out_t = torch.empty((count, n_snap), dtype=tdt, device=dev)
            out_x = torch.empty((count, n_snap, d), dtype=tdt, device=dev)
            out_w = torch.empty((count, n_snap, m), dtype=tdt, device=dev)
        moments = None
        if n_obs:
            shape = (n_snap, n_obs, 2) if observe_at == "snapshots" else (n_obs, 2)
            moments = torch.zeros(shape, dtype=torch.float64, device=dev)
        out_hit = torch.empty((count, 2 + d), dtype=tdt, device=dev) if first_passage is not None else None

        ka, kb = prng_key(seed, x64=f64 or config.get("enable_x64"))
        args = runtime.SolveArgs(
            key_a=ka, key_b=kb, n_fragments=n_frag, cpr=cpr, chunk=chunk, n_snapshots=n_snap,
            n_traj_total=n_total, traj_begin=begin, traj_count=count,
            x0=x0_dev.data_ptr(), x0_stride=(d if multiple_ic else 0),
            dt=float(self.dt), dt32=float(self.dt) ** (3 / 2),
            out_t=out_t.data_ptr() if out_t is not None else None,
            out_x=out_x.data_ptr() if out_x is not None else None,
            out_w=out_w.data_ptr() if out_w is not None else None,
            moments=moments.data_ptr() if moments is not None else None,
            out_hit=out_hit.data_ptr() if out_hit is not None else None)
        if count > 0:
            runtime.check(runtime.lib().sdeb200_solve(prog.handle, ctypes.byref(args), block, traj_per_block, dyn_smem,
                                                      runtime.current_stream_ptr()), "sdeb200_solve")
            runtime.STATS["launches"] += 1
        self.last_program = prog
        self.last_steps = n_snap * chunk

        result = {}
        if store_trajectories:
            result = dict(time_values=DeviceArray(out_t), solution_values=DeviceArray(out_x),
                          wiener_values=DeviceArray(out_w))
        if moments is not None:
            result["moments"] = DeviceArray(moments)
            result["count"] = count
        if out_hit is not None:
            result["hit_flag"] = DeviceArray(out_hit[:, 0] > 0.5)
            result["hit_time"] = DeviceArray(out_hit[:, 1])
            result["hit_state"] = DeviceArray(out_hit[:, 2:])
        sol = _Solution(result)
        sol._keepalive = (x0_dev,)        # the launch is asynchronous: its inputs live as long as its results
        return sol

    def _solve_bead_ring(self, problem, n_trajectories, seed, chunk_size, chunks_per_randomization, traj_range,
                         max_snapshots, prng_layout):
        """Euler-Maruyama for `workloads.BeadRingProblem` on the CTA-per-trajectory kernel
        (csrc/sde_bead_kernel.cuh): mobility, Cholesky factor and forces in shared memory."""
        if self.scheme != "euler":
            raise NotImplementedError("bead chains with Cholesky noise are integrated with the Euler scheme "
                                      "(as in every bead example of the reference)")
        x0 = np.asarray(jnp._host(problem.x0))
        multiple_ic = x0.ndim == 2
        if multiple_ic and n_trajectories is not None:
            raise ValueError("n_trajectories option not supproted with multiple inital conditions!")
        f64 = x0.dtype == np.float64
        x0 = x0.astype(np.float64 if f64 else np.float32)
        ics = x0 if multiple_ic else x0.reshape(1, -1)
        n_total = ics.shape[0] if multiple_ic else int(n_trajectories)
        par = problem.bead_ring
        nb = par["n_beads"]
        d = 3 * nb
        layout = prng_layout or config.get("prng_layout")
        plan = self._bead_plan(problem, f64, layout)
        block = plan.ks.block
        prog = plan.program()
        torch = runtime.require_cuda()
        chunk, n_chunks, cpr, n_frag = chunk_arithmetic(problem.tmax, self.dt, chunk_size, chunks_per_randomization)
        n_snap = n_chunks if max_snapshots is None else min(n_chunks, int(max_snapshots))
        if 2 * chunk * cpr * d >= 2 ** 32:
            raise ValueError("chunk_size * chunks_per_randomization too large for the 32-bit counter layout")
        begin, count = (0, n_total) if traj_range is None else (int(traj_range[0]), int(traj_range[1]))
        dev = torch.device("cuda", torch.cuda.current_device())
        tdt = torch.float64 if f64 else torch.float32
        ic_host = ics[begin:begin + count] if multiple_ic else ics
        x0_dev = torch.as_tensor(np.ascontiguousarray(ic_host), dtype=tdt).to(dev, non_blocking=True)
        out_t = torch.empty((count, n_snap), dtype=tdt, device=dev)
        out_x = torch.empty((count, n_snap, d), dtype=tdt, device=dev)
        out_w = torch.empty((count, n_snap, d), dtype=tdt, device=dev)
        ka, kb = prng_key(seed, x64=f64 or config.get("enable_x64"))
        args = runtime.BeadArgs(
            key_a=ka, key_b=kb, n_fragments=n_frag, cpr=cpr, chunk=chunk, n_snapshots=n_snap, n_traj_total=n_total,
            traj_begin=begin, traj_count=count, x0=x0_dev.data_ptr(), x0_stride=(d if multiple_ic else 0),
            dt=float(self.dt), radius=par["radius"], ks=par["ks"], d0=par["d0"], kb_over_d04=par["kb_over_d04"],
            ko=par["ko"], sigma2=par["sigma2"], inv_delta2=par["inv_delta2"], twist_k=par.get("twist_k", 0.0),
            linking_number=par.get("linking_number", 0.0), workspace=None,
            out_t=out_t.data_ptr(), out_x=out_x.data_ptr(), out_w=out_w.data_ptr())
        elem = 8 if f64 else 4
        workspace = None
        if plan.ks.traj_per_block == 2:            # register-tiled variant (traj_per_block doubles as its marker)
            smem = bead_regtile_smem_bytes(nb, bool(par.get("twist")))
            ctas_per_sm = max(1, min(2 if block > 170 else 3, (227 * 1024) // (smem + 1024)))
            npad = -(-d // 8) * 8
            # value-table indices of every (tile, lane), then 1024 CTA arrival counters per SM (must be zero)
            workspace = torch.zeros(((npad // 8) * (npad // 8 + 1) // 2) * 32 + 1024, dtype=torch.int32, device=dev)
            args.workspace = workspace.data_ptr()
        else:
            npad = -(-d // 8) * 8                  # must match SDE_NP / SDE_LSIZE in csrc/sde_bead_kernel.cuh
            lsize = 17 * ((npad // 8) * npad - 4 * (npad // 8) * (npad // 8 - 1)) // 2 + 2
            smem = (7 * d + 8 + lsize) * elem
            ctas_per_sm = max(1, min(8, (220 * 1024) // (smem + 1024)))
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        runtime.check(runtime.lib().sdeb200_bead_solve(prog.handle, ctypes.byref(args), block, smem,
                                                       sms * ctas_per_sm, runtime.current_stream_ptr()),
                      "sdeb200_bead_solve")
        runtime.STATS["launches"] += 1
        self.last_program = prog
        self.last_steps = n_snap * chunk
        sol = _Solution(dict(time_values=DeviceArray(out_t), solution_values=DeviceArray(out_x),
                             wiener_values=DeviceArray(out_w)))
        sol._keepalive = (x0_dev, workspace)
        return sol

    def solve(self, problem, seed=0, chunk_size=1, chunks_per_randomization=None, progress_bar=True):
        """One trajectory, leading axis stripped (reference `sde_solver.py:306-359`)."""
        sol = self.solve_many(problem, n_trajectories=1, seed=seed, chunk_size=chunk_size,
                              chunks_per_randomization=chunks_per_randomization, progress_bar=progress_bar)
        return _Solution({k: (v[0] if isinstance(v, DeviceArray) else v) for k, v in sol.items()})


class _Solution(dict):
    """Result dict; hides bookkeeping entries from iteration over the reference's three keys."""

    def block_until_ready(self):
        for v in self.values():
            if isinstance(v, DeviceArray):
                v.block_until_ready()
        return self
