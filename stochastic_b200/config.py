"""Package-wide switches (the counterpart of `jax.config`)."""

_STATE = {
    # Mirrors `jax_enable_x64`: when True, SDEProblem keeps x0 in float64 (the reference casts x0 to
    # float32 unconditionally, reference pychastic/sde_problem.py:51, and needs x0 overwritten after
    # construction to run in double precision).
    "enable_x64": False,
    # "original": threefry counter layout of jax < 0.5 (jax_threefry_partitionable=False; the layout the
    # reference's shipped golden trajectories were generated with); "partitionable": jax >= 0.5 default.
    "prng_layout": "original",
    # Wagner-Platen with several noise terms: use the producer/consumer kernel (csrc/sde_kernel_ws.cuh) whenever
    # its two-stage ring of staged normals fits in shared memory.  False forces the single-role kernel.
    "warp_specialize": True,
    # Emitter: a denominator that divides two or more expressions of a step function is inverted once and the
    # divisions become multiplications (an FP64 division costs ~12 instructions on the GPU).  Changes the last bit of
    # those quotients, so it is opt-in: the default keeps the operations the user's functions spell out.
    "reciprocal_division": False,
    # Library routines on traced fp64 input take every inverse power of a quantity from ONE reciprocal square root (the CUDA
    # math library's 1-ulp `rsqrt`, ~12 instructions) instead of square roots and divisions (~25-50 each):
    # `jnp.linalg.cholesky` (L_jj = s rsqrt(s), column x rsqrt(s)) and `grpy.muTT` (1/r, 1/r^2, 1/r^3 of a pair).  How a
    # library routine forms its result is the library's choice - XLA's Cholesky expander and pygrpy make their own - so
    # this is not covered by "reciprocal_division", which is about the user's own arithmetic.  Results move by ~1e-15;
    # two spheres / four beads gain 4-5 % from the factor alone.  False restores sqrt + division; fp32 always uses those.
    "library_rsqrt": True,
    # Stored trajectories: hand whole snapshot tiles (32 trajectories x 16..128 bytes per field column) to the TMA
    # unit (csrc/sde_snap_tma.cuh) whenever the output rows are 16-byte multiples.  False: the warp-transposed
    # write-out loop of csrc/sde_kernel.cuh (also the fallback for rows a tensor map cannot describe).
    "snapshot_tma": True,
}


def update(name, value):
    name = name.replace("jax_", "")
    if name == "threefry_partitionable":
        name, value = "prng_layout", ("partitionable" if value else "original")
    if name not in _STATE:
        raise KeyError(name)
    if name == "prng_layout" and value not in ("original", "partitionable"):
        raise ValueError(value)
    _STATE[name] = value


def get(name):
    return _STATE[name]
