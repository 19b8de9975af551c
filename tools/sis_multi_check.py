"""SIS-aCS and CMC sharded over G GPUs: estimates against the truth (statistical check; the multinomial split of the
seeds over the ranks makes multi-GPU runs a different, equally distributed stream).
torchrun --nproc-per-node 2 tools/sis_multi_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rareeventestimation_b200 as ree  # noqa: E402
from rareeventestimation_b200.engine import Comm, Engine  # noqa: E402
from rareeventestimation_b200.problem import DeviceSample  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine.get(torch.device("cuda", local), Comm())
errs = []
for d, n in [(2, 20_000), (100, 200_000)]:
    prob = ree.make_linear_problem(d)
    est = []
    for s in range(6):
        prob.sample = DeviceSample(n, d, seed=100 + s)
        sol = ree.SIS(seed=s, mixture_model="aCS").solve(prob)
        est.append(float(sol.prob_fail_hist[-1]))
    est = np.array(est)
    rel = np.abs(est / prob.prob_fail_true - 1)
    if rank == 0:
        print(f"SIS-aCS d={d} N={n} world={world}: estimates/truth {np.round(est / prob.prob_fail_true, 3)} "
              f"levels={sol.num_steps} costs={sol.costs}", flush=True)
    errs.append(rel.mean())
    assert rel.max() < 0.5 and abs(est.mean() / prob.prob_fail_true - 1) < 0.15, (d, est)
prob = ree.make_linear_problem(100)
sol = ree.CMC(20_000_000, seed=2, verbose=False).solve(prob)
pf = float(sol.prob_fail_hist[-1])
if rank == 0:
    print(f"CMC d=100 N=2e7 world={world}: pf/truth {pf / prob.prob_fail_true:.4f}", flush=True)
assert abs(pf / prob.prob_fail_true - 1) < 0.05
if rank == 0:
    print("OK", flush=True)
if world > 1:
    dist.destroy_process_group()
