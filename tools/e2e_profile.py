"""Phases of the end-to-end call CBREE(...).solve(problem) with the sample in pinned host memory (bench.py's e2e leg).
Usage: python tools/e2e_profile.py [--n 10000000] [--d 100] [--steps 10]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rareeventestimation_b200 as ree  # noqa: E402
from rareeventestimation_b200.engine import Engine  # noqa: E402
from rareeventestimation_b200.problem import DeviceSample, ShardedHostSample  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--d", type=int, default=100)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
eng = Engine.get()
n, d = args.n, args.d
host = torch.empty((n, d), dtype=torch.float64, pin_memory=True)
src = DeviceSample(n, d, seed=0x5EED).materialise(eng, 0, n)
eng.download(src, out=host.numpy())
del src
torch.cuda.empty_cache()


def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"  {label:34s} {1e3 * (t1 - t0):9.2f} ms", flush=True)
    return t1


for rep in range(2):
    prob = ree.make_linear_problem(d)
    prob.sample = ShardedHostSample(host.numpy(), n, 0)
    solver = ree.CBREE(num_steps=args.steps, seed=1, convergence_check=False, divergence_check=False, quiet=True)
    print(f"rep {rep}")
    torch.cuda.synchronize()
    t00 = t0 = time.perf_counter()
    cl = solver.start(prob, recycle=True)
    t0 = tick("start (upload, g/e, weights, h0)", t0)
    for i in range(args.steps):
        solver.advance(cl)
    t0 = tick(f"{args.steps} x advance", t0)
    sol = solver.finish(cl)
    t0 = tick("finish (Solution assembly)", t0)
    pf = float(sol.prob_fail_hist[-1])
    g_last = np.asarray(sol.lsf_eval_hist[-1])
    t0 = tick("read prob_fail_hist, lsf_eval_hist[-1]", t0)
    print(f"  total {1e3 * (t0 - t00):.2f} ms  pf={pf:.4e}")
    del sol, solver, cl, prob
    torch.cuda.empty_cache()
