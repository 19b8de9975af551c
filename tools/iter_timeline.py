"""Device timeline of single CBREE iterations: start offset, duration and the idle gap before every engine call
(CUDA events on the launching stream).  Usage: python tools/iter_timeline.py --n 1250000 [--iters 2]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rareeventestimation_b200 as ree  # noqa: E402
from rareeventestimation_b200.engine import Engine  # noqa: E402
from rareeventestimation_b200.problem import DeviceSample  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_250_000)
ap.add_argument("--d", type=int, default=100)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--problem", default="linear", choices=["linear", "osc"])
ap.add_argument("--skip", type=int, default=12, help="iterations before the recorded ones (the Brent search is active after ~9)")
args = ap.parse_args()
eng = Engine.get()
events = []
NAMES = ["cbree_step", "cbree_moments", "cbree_maha", "reduce_ess", "moments_chol", "chol_scaled", "to_device", "adapt",
         "fetch", "fetch_array", "ensemble_sums", "count_failed"]
for name in NAMES:
    fn = getattr(eng, name)

    def wrap(fn=fn, name=name):
        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            out = fn(*a, **k)
            e1.record()
            events.append((name, e0, e1, t0, time.perf_counter()))
            return out
        return inner
    setattr(eng, name, wrap())

if args.problem == "osc":
    prob, args.d = ree.prob_nonlin_osc, 6
else:
    prob = ree.make_linear_problem(args.d)
prob.sample = DeviceSample(args.n, args.d, seed=0x5EED)
solver = ree.CBREE(seed=1, convergence_check=False, divergence_check=False, quiet=True, num_steps=10**9)
cl = solver.start(prob, recycle=True)
for _ in range(args.skip):
    solver.advance(cl)
torch.cuda.synchronize()
for it in range(args.iters):
    events.clear()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    s0.record()
    solver.advance(cl)
    s1.record()
    torch.cuda.synchronize()
    h1 = time.perf_counter()
    print(f"iteration {cl[-1].iteration}: device {s0.elapsed_time(s1):.3f} ms, host {1e3 * (h1 - h0):.3f} ms   (n={args.n}, d={args.d})")
    prev_end = 0.0
    busy = 0.0
    for name, e0, e1, t0, t1 in events:
        st, en = s0.elapsed_time(e0), s0.elapsed_time(e1)
        busy += en - st
        print(f"  {name:14s} dev start {st:8.3f}  dur {en - st:7.3f}  gap before {st - prev_end:7.3f} | host call at {1e3 * (t0 - h0):7.3f} "
              f"took {1e3 * (t1 - t0):6.3f}")
        prev_end = en
    print(f"  tail gap {s0.elapsed_time(s1) - prev_end:.3f} ms; sum of call durations {busy:.3f} ms")
