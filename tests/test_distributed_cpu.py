"""world_size-2 gloo tests of the multi-GPU host logic (sharding + allreduce of moments), on the CPU.
The oracle stands in for the kernel: the union of the shards must reproduce the unsharded run."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stochastic_b200.distributed import allreduce_moments, mean_and_stderr, shard_range


def test_shard_range_covers_everything():
    for n, w in [(10, 1), (10, 2), (10, 3), (10 ** 9, 8), (7, 8), (0, 4)]:
        parts = [shard_range(n, r, w) for r in range(w)]
        assert sum(c for _, c in parts) == n
        pos = 0
        for b, c in parts:
            assert b == pos or c == 0
            pos = b + c


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import problems, solver
    prob = problems.garch(tmax=0.25)
    begin, count = shard_range(n_total, rank, world)
    sol = solver.solve_many(prob, scheme="milstein", dt=2 ** -4, n_trajectories=n_total, seed=7, chunk_size=None,
                            dtype=np.float64, traj_slice=(begin, count))
    payoff = np.maximum(sol["solution_values"][:, -1, 0] - 100.0, 0.0)
    mom = torch.tensor([[payoff.sum(), (payoff ** 2).sum()]], dtype=torch.float64)
    mom, total = allreduce_moments(mom, count)
    mean, se = mean_and_stderr(mom, total)
    if rank == 0:
        out.put((mom.numpy().copy(), total, float(mean[0]), float(se[0])))
    dist.destroy_process_group()


def test_sharded_moments_equal_unsharded_moments():
    from oracle import problems, solver
    n_total, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, PORT, n_total, q))
             for r in range(world)]
    for p in procs:
        p.start()
    mom, total, mean, se = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sol = solver.solve_many(problems.garch(tmax=0.25), scheme="milstein", dt=2 ** -4, n_trajectories=n_total, seed=7,
                            chunk_size=None, dtype=np.float64)
    payoff = np.maximum(sol["solution_values"][:, -1, 0] - 100.0, 0.0)
    assert total == n_total
    assert np.allclose(mom[0], [payoff.sum(), (payoff ** 2).sum()], rtol=1e-13)
    assert np.isclose(mean, payoff.mean())


PORT = _free_port()
