"""Where does one CBREE iteration spend its time?  Wraps every Engine method with CUDA events (device time) and
host timers; prints per-method totals for K iterations at (n, d).  Usage: python tools/iter_profile.py --n 10000000"""
import argparse
import os
import sys
import time
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rareeventestimation_b200 as ree  # noqa: E402
from rareeventestimation_b200.engine import Engine  # noqa: E402
from rareeventestimation_b200.problem import DeviceSample  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--d", type=int, default=100)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--problem", default="linear", choices=["linear", "osc", "pde"])
ap.add_argument("--resample", action="store_true", help="vMFNM resampling branch (resample=True, mixture_model='vMFNM')")
args = ap.parse_args()
eng = Engine.get()
dev_ms, host_ms, calls = defaultdict(float), defaultdict(float), defaultdict(int)
events = []
for name in ["cbree_step", "cbree_moments", "cbree_maha", "reduce_ess", "reduce_cvar", "reduce_scaled", "logtgt", "moments_chol",
             "chol_scaled", "ensemble_sums", "count_failed", "to_device", "alloc_ensemble", "empty", "lsf_eval", "energy",
             "vmfn_fit_sums", "vmfn_sample", "vmfn_logpdf", "vector_range", "adapt"]:
    fn = getattr(eng, name)

    def wrap(fn=fn, name=name):
        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            out = fn(*a, **k)
            e1.record()
            host_ms[name] += (time.perf_counter() - t0) * 1e3
            events.append((name, e0, e1))
            calls[name] += 1
            return out
        return inner
    setattr(eng, name, wrap())

if args.problem == "osc":
    prob, args.d = ree.prob_nonlin_osc, 6
elif args.problem == "pde":
    prob, args.d = ree.diffusion_problem, 150
else:
    prob = ree.make_linear_problem(args.d)
prob.sample = DeviceSample(args.n, args.d, seed=0x5EED)
extra = dict(resample=True, mixture_model="vMFNM") if args.resample else {}
solver = ree.CBREE(seed=1, convergence_check=False, divergence_check=False, quiet=True, num_steps=10**9, **extra)
cl = solver.start(prob, recycle=True)
for _ in range(3):
    solver.advance(cl)
torch.cuda.synchronize()
events.clear(); dev_ms.clear(); host_ms.clear(); calls.clear()
t0 = time.perf_counter()
for _ in range(args.steps):
    solver.advance(cl)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
for name, e0, e1 in events:
    dev_ms[name] += e0.elapsed_time(e1)
print(f"wall per iteration: {wall / args.steps:.2f} ms  (n={args.n}, d={args.d})")
tot = 0.0
for name in sorted(dev_ms, key=lambda k: -dev_ms[k]):
    print(f"  {name:16s} calls/it {calls[name] / args.steps:6.1f}  device {dev_ms[name] / args.steps:8.3f} ms/it  "
          f"host {host_ms[name] / args.steps:8.3f} ms/it")
    tot += dev_ms[name]
print(f"  sum of device time {tot / args.steps:.2f} ms/it")
