"""Timings of the other BASELINE.json configurations on one GPU (one rank's share of the multi-GPU ones).
Usage: python tools/config_bench.py [pde] [osc] [cmc] [sis]     -> one JSON line per configuration."""
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
eng = Engine.get()


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
    ms = e0.elapsed_time(e1) / steps
    return ms, t_start, cl[-1]


if "pde" in which:
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
    print(json.dumps({"config": "CBREE diffusion PDE d=150 N=1e6", "ms_per_iteration": ms, "p_it_per_s": n / ms * 1e3,
                      "lsf_ms_per_sweep": lsf_ms, "lsf_evals_per_s": n / lsf_ms * 1e3, "lsf_fp64_tflops": flop / lsf_ms / 1e9,
                      "start_s": t_start, "sigma": last.sigma, "path": eng.path(d)}), flush=True)
    eng.drop_pool()

if "osc" in which:
    n, d = 12_500_000, 6  # one rank's share of N=1e8 over 8 GPUs
    ms, t_start, last = cbree_iterations(ree.prob_nonlin_osc, n, d)
    bytes_it = n * (24 * d + 48)
    print(json.dumps({"config": "CBREE oscillator d=6, N=1.25e7 (1/8 of 1e8)", "ms_per_iteration": ms,
                      "p_it_per_s": n / ms * 1e3, "hbm_frac_iteration": bytes_it / ms / 1e6 / HBM, "start_s": t_start,
                      "path": eng.path(d)}), flush=True)
    eng.drop_pool()

if "cmc" in which:
    n, d = 100_000_000, 100
    prob = ree.make_linear_problem(d)
    ree.CMC(4_000_000, seed=1, verbose=False).solve(prob)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sol = ree.CMC(n, seed=2, verbose=False).solve(prob)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "CMC affine d=100, N=1e8", "seconds": dt, "samples_per_s": n / dt,
                      "pf": float(sol.prob_fail_hist[-1]), "truth": float(prob.prob_fail_true)}), flush=True)
    eng.drop_pool()

if "sis" in which:
    n, d = 10_000_000, 100
    prob = ree.make_linear_problem(d)
    prob.sample = DeviceSample(n, d, seed=5)
    t0 = time.perf_counter()
    sol = ree.SIS(seed=1, mixture_model="aCS").solve(prob)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"config": "SIS-aCS affine d=100, N=1e7 per level", "seconds": dt, "lsf_evals": int(sol.costs),
                      "lsf_evals_per_s": sol.costs / dt, "levels": int(sol.num_steps),
                      "pf": float(sol.prob_fail_hist[-1]), "truth": float(prob.prob_fail_true)}), flush=True)
