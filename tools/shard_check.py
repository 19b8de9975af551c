"""Sharding invariance of a whole solve: the same global problem on 1 and on G GPUs must give the same trajectory
(Philox counters are global particle indices; only the order of the cross-rank sums differs).

  python tools/shard_check.py --out gpurun_out/shard_1.json
  torchrun --nproc-per-node 2 tools/shard_check.py --out gpurun_out/shard_2.json
  python tools/shard_check.py --compare gpurun_out/shard_1.json gpurun_out/shard_2.json
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--out")
ap.add_argument("--compare", nargs=2)
ap.add_argument("--particles", type=int, default=200_000)
args = ap.parse_args()

if args.compare:
    a, b = (json.load(open(f)) for f in args.compare)
    worst = 0.0
    for cfg in a:
        for key in ("sigma", "beta", "t_step", "cvar", "estimate"):
            x, y = np.array(a[cfg][key], dtype=float), np.array(b[cfg][key], dtype=float)
            assert x.shape == y.shape, (cfg, key, x.shape, y.shape)
            err = np.nanmax(np.abs(x - y) / np.maximum(np.abs(x), 1e-300))
            worst = max(worst, err)
            print(f"{cfg:28s} {key:9s} max rel diff {err:.2e}")
        assert a[cfg]["costs"] == b[cfg]["costs"] and a[cfg]["msg"] == b[cfg]["msg"]
    assert worst < 1e-6, worst
    print("OK: trajectories agree, worst relative difference", worst)
    sys.exit(0)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import rareeventestimation_b200 as ree  # noqa: E402
from rareeventestimation_b200.engine import Comm, Engine  # noqa: E402
from rareeventestimation_b200.problem import DeviceSample  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine.get(torch.device("cuda", local), Comm())
out = {}
for name, prob, d, kw in [("linear d=100", ree.make_linear_problem(100), 100, {}),
                          ("oscillator d=6", ree.prob_nonlin_osc, 6, {}),
                          ("linear d=20 vMFNM resample", ree.make_linear_problem(20), 20,
                           dict(resample=True, mixture_model="vMFNM"))]:
    prob.sample = DeviceSample(args.particles, d, seed=0x5EED)
    sol = ree.CBREE(seed=3, num_steps=8, convergence_check=False, divergence_check=False, quiet=True,
                    return_caches=True, save_history=True, **kw).solve(prob)
    caches = sol.other["cache_list"]
    out[name] = {"sigma": [c.sigma for c in caches], "beta": [c.beta for c in caches],
                 "t_step": [float(c.t_step) for c in caches], "cvar": [float(c.cvar_is_weights) for c in caches],
                 "estimate": [float(x) for x in sol.prob_fail_hist], "costs": int(sol.costs), "msg": sol.msg}
    eng.drop_pool()
if rank == 0:
    json.dump(out, open(args.out, "w"))
    print("wrote", args.out, {k: v["estimate"][-1] for k, v in out.items()})
if world > 1:
    dist.destroy_process_group()
