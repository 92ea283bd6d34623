"""Multi-GPU: shard trajectories over ranks, allreduce the moments.

The reference has no multi-device code at all.  Trajectories are independent, so
rank r of R integrates the contiguous slice [r*ceil(N/R), min(N, (r+1)*ceil(N/R))) of the GLOBAL
key fan-out `split(PRNGKey(seed), N)` - every rank derives its own keys in O(1) per trajectory inside
the kernel, nothing is broadcast, and the union of the shards is bit-identical to a single-process
`solve_many(n_trajectories=N)`.  The only exchange is one allreduce(sum) of the tiny fp64 moments
vector (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).
"""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world_size):
    """(begin, count) of rank's slice; slices are contiguous, disjoint and cover [0, n_total)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    per = -(-int(n_total) // int(world_size))
    begin = min(int(n_total), rank * per)
    end = min(int(n_total), begin + per)
    return begin, end - begin


def allreduce_moments(moments, count, group=None):
    """Sum a (n_obs, 2) moments tensor and the trajectory count over all ranks (in place for moments).

    Returns (moments, total_count).  No-op outside an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return moments, int(count)
    packed = torch.cat([moments.reshape(-1).to(torch.float64),
                        torch.tensor([float(count)], dtype=torch.float64, device=moments.device)])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    moments.copy_(packed[:-1].reshape(moments.shape))
    return moments, int(round(float(packed[-1].item())))


def mean_and_stderr(moments, count):
    """Per-observable mean and standard error of the mean from (sum, sum of squares)."""
    m = moments.to(torch.float64)
    mean = m[:, 0] / count
    var = torch.clamp(m[:, 1] / count - mean * mean, min=0.0)
    return mean, torch.sqrt(var / count)


def solve_many_sharded(solver, problem, n_trajectories, observables=None, group=None, key_mode="global", seed=0,
                       **kwargs):
    """`solver.solve_many` on this rank's shard of an n_trajectories-wide run.

    key_mode="global" (default): every rank uses its slice of the GLOBAL fan-out split(PRNGKey(seed), N) - the
    union of the shards is bit-identical to a single-process solve_many(n_trajectories=N).
    key_mode="fold_in": rank r runs an independent count_r-trajectory solve keyed by
    fold_in(PRNGKey(seed), r) - the scheme BASELINE.json's north_star names; a different (equally valid) stream,
    with no limit on the global trajectory count.

    Returns the rank-local solution dict; if `observables` is given its 'moments' / 'count' entries hold the
    GLOBAL sums after one allreduce."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    begin, count = shard_range(n_trajectories, rank, world)
    if key_mode == "global":
        sol = solver.solve_many(problem, n_trajectories=n_trajectories, observables=observables, seed=seed,
                                traj_range=(begin, count), **kwargs)
    elif key_mode == "fold_in":
        from . import prng
        sol = solver.solve_many(problem, n_trajectories=count, observables=observables,
                                seed=prng.fold_in(prng.PRNGKey(seed), rank), **kwargs)
    else:
        raise ValueError(f"unknown key_mode {key_mode!r}")
    sol["traj_range"] = (begin, count)
    if observables is not None:
        mom, total = allreduce_moments(sol["moments"].torch(), sol["count"], group)
        sol["count"] = total
    return sol
