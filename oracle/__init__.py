"""CPU oracle for the batched SDE-integration hot path (TEST INFRASTRUCTURE ONLY).

This package is a numpy/torch-CPU restatement of the algorithm of the reference
(`pychastic.sde_solver.SDESolver.solve_many`, reference `pychastic/sde_solver.py:82-304`)
and of the slice of its un-vendored dependency `jax`/`jaxlib` (pinned `jax==0.4.35`,
reference `docs/requirements.txt:16-17`) that decides the random stream.

It exists to CHECK the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product
package `stochastic_b200` never imports anything from here and has no CPU fallback.

Parity status
-------------
* fp32 stream + Euler stepping + RPY mobility: PINNED by the reference's shipped golden
  trajectories (`tests/golden/stopping_time_trajectory_*.csv`,
  `tests/golden/bead_single_trajectory_head.csv`; see `tests/test_oracle_golden.py`).
* threefry2x32 / key splitting: PINNED by Random123 and JAX known-answer values.
* fp64 stream (64 random bits per draw): the BITS are pinned by jax's own regression vectors
  for `random_bits(key(1701), 64, (3,))` (jax tests/random_test.py::testRngRandomBits, with and
  without x64); the map bits -> uniform -> normal shares its code with the pinned fp32 twin
  (52 instead of 23 mantissa bits, XLA's f64 `erf_inv` polynomial checked against scipy).  No
  fp64 TRAJECTORY fixture exists - JAX cannot be installed here.
* Wagner-Platen stepping: pinned only statistically (iterated-integral moments against
  `wiener_integral_moments.E/E2`, strong/weak orders), as in the reference's own tests.
"""
