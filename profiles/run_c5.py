"""BASELINE config C5: GARCH / "Heston-like" stochastic-volatility call pricing (reference
docs/source/optionpricing.rst:117-167), Wagner-Platen fp64, paths sharded over the GPUs of one box, one NCCL
allreduce of the payoff moments.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 profiles/run_c5.py [n_paths]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import stochastic_b200 as pychastic
from stochastic_b200 import workloads
from stochastic_b200.distributed import mean_and_stderr, solve_many_sharded

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
n_paths = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 9
problem = workloads.garch(f64=True)
solver = pychastic.sde_solver.SDESolver(scheme="wagner_platen", dt=0.01)
payoff = workloads.call_payoff(120.0)


def run(n):
    return solve_many_sharded(solver, problem, n, observables=payoff, seed=0, chunk_size=None, progress_bar=False,
                              store_trajectories=False)


run(world * 10 ** 5)                      # warm-up (NVRTC / cache, NCCL)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
sol = run(n_paths)
ev1.record()
torch.cuda.synchronize()
ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
mean, se = mean_and_stderr(sol["moments"].torch(), sol["count"])
if rank == 0:
    steps = solver.last_steps
    out = {"config": "C5 GARCH call payoff (strike 120), wagner_platen fp64, moments only", "n_gpus": world,
           "n_paths": sol["count"], "steps": steps, "ms": float(ms.item()),
           "traj_steps_per_s": sol["count"] * steps / float(ms.item()) * 1e3,
           "price": float(mean[0]), "stderr": float(se[0])}
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/c5_n{world}.json", "w"))
if world > 1:
    dist.destroy_process_group()
