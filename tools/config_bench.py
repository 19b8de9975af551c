"""Timings of the other BASELINE.json configurations.
  one GPU (one rank's share of the multi-GPU configurations):
      python tools/config_bench.py [pde] [osc] [cmc] [sis]
  the configurations BASELINE.json quotes on 8 GPUs, at their full size, particles sharded over the ranks:
      python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config_bench.py osc cmc sis
      (configs[3]: CBREE oscillator d=6, N=1e8; configs[4]: CMC and SIS-aCS, affine d=100, N=1e8)
-> one JSON line per configuration (rank 0); times are device times, the maximum over the ranks."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rareeventestimation_b200 as ree  # noqa: E402
from rareeventestimation_b200.engine import Engine  # noqa: E402
from rareeventestimation_b200.problem import DeviceSample  # noqa: E402

HBM = 6546.6
which = set(sys.argv[1:]) or {"pde", "osc", "cmc", "sis"}
world = int(os.environ.get("WORLD_SIZE", "1"))
rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    from rareeventestimation_b200.engine import Comm

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine.get(torch.device("cuda", local), Comm())
else:
    eng = Engine.get()


def sync_max(x):
    """max over the ranks of a host number (after a device synchronisation)."""
    torch.cuda.synchronize()
    t = torch.tensor([x], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def say(d):
    if rank == 0:
        d["n_gpus"] = world
        print(json.dumps(d), flush=True)



def cbree_iterations(prob, n, d, steps=6, warm=2, **kw):
    prob.sample = DeviceSample(n, d, seed=0x5EED)
    solver = ree.CBREE(seed=1, convergence_check=False, divergence_check=False, quiet=True, num_steps=10**9, **kw)
    t0 = time.perf_counter()
    cl = solver.start(prob, recycle=True)
    torch.cuda.synchronize()
    t_start = time.perf_counter() - t0
    for _ in range(warm):
        solver.advance(cl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        solver.advance(cl)
    e1.record()
    torch.cuda.synchronize()
    ms = sync_max(e0.elapsed_time(e1)) / steps
    return ms, t_start, cl[-1]


if "pde" in which and world == 1:
    n, d = 1_000_000, 150
    prob = ree.diffusion_problem
    ens = eng.philox_normal(n, d, 1, 1)
    desc = prob.lsf.descriptor(eng, d)
    eng.lsf_eval(desc, ens, out=ens.g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.lsf_eval(desc, ens, out=ens.g)
    e1.record()
    torch.cuda.synchronize()
    lsf_ms = e0.elapsed_time(e1) / 3
    flop = n * (2.0 * 1536 * d + 1536 * 30 + 512 * 8)
    del ens
    eng.drop_pool()
    ms, t_start, last = cbree_iterations(prob, n, d, steps=4, warm=1)
    say({"config": "CBREE diffusion PDE d=150 N=1e6", "ms_per_iteration": ms, "p_it_per_s": n / ms * 1e3,
         "lsf_ms_per_sweep": lsf_ms, "lsf_evals_per_s": n / lsf_ms * 1e3, "lsf_fp64_tflops": flop / lsf_ms / 1e9,
         "start_s": t_start, "sigma": last.sigma, "path": eng.path(d)})
    eng.drop_pool()

if "osc" in which:
    n, d = (100_000_000, 6) if world > 1 else (12_500_000, 6)  # N=1e8 over the ranks; alone: one rank's share of 8
    ms, t_start, last = cbree_iterations(ree.prob_nonlin_osc, n, d)
    bytes_it = n / world * (24 * d + 48)
    say({"config": f"CBREE oscillator d=6, N={n:.3g}" + ("" if world > 1 else " (1/8 of 1e8)"), "ms_per_iteration": ms,
         "p_it_per_s": n / ms * 1e3, "hbm_frac_iteration": bytes_it / ms / 1e6 / HBM, "start_s": t_start,
         "path": eng.path(d)})
    eng.drop_pool()
    # the complete solve with the reference's defaults (test/test_nonlinear_oscillator_prob.py:4-9 at this size)
    prob = ree.prob_nonlin_osc
    prob.sample = DeviceSample(n, d, seed=0x5EED)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sol = ree.CBREE(seed=1, quiet=True).solve(prob)
    dt = sync_max(time.perf_counter() - t0)
    say({"config": f"CBREE oscillator d=6, N={n:.3g}: complete solve, reference defaults", "seconds": dt,
         "steps": int(sol.num_steps), "msg": sol.msg, "pf": float(sol.prob_fail_hist[-1]), "truth": float(prob.prob_fail_true),
         "rel_err": abs(float(sol.prob_fail_hist[-1]) / float(prob.prob_fail_true) - 1), "lsf_evals": int(sol.costs),
         "p_it_per_s": sol.costs / dt})
    del sol
    eng.drop_pool()

if "cmc" in which:
    n, d = 100_000_000, 100
    prob = ree.make_linear_problem(d)
    ree.CMC(4_000_000, seed=1, verbose=False).solve(prob)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sol = ree.CMC(n, seed=2, verbose=False).solve(prob)
    dt = sync_max(time.perf_counter() - t0)
    say({"config": "CMC affine d=100, N=1e8", "seconds": dt, "samples_per_s": n / dt,
         "pf": float(sol.prob_fail_hist[-1]), "truth": float(prob.prob_fail_true)})
    eng.drop_pool()

if "sis" in which:
    n, d = (100_000_000, 100) if world > 1 else (10_000_000, 100)  # population per level (80 GB at 1e8: sharded)
    prob = ree.make_linear_problem(d)
    prob.sample = DeviceSample(n, d, seed=5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sol = ree.SIS(seed=1, mixture_model="aCS").solve(prob)
    dt = sync_max(time.perf_counter() - t0)
    say({"config": f"SIS-aCS affine d=100, N={n:.3g} per level", "seconds": dt, "lsf_evals": int(sol.costs),
         "lsf_evals_per_s": sol.costs / dt, "levels": int(sol.num_steps),
         "pf": float(sol.prob_fail_hist[-1]), "truth": float(prob.prob_fail_true)})
if world > 1:
    dist.destroy_process_group()
