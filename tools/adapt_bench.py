"""Timing of the persistent sigma / beta search kernel (ree_cbree_adapt) alone: per-evaluation cost vs particle count.
Usage: python tools/adapt_bench.py [--reps 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rareeventestimation_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--sizes", default="1000,156250,1250000,2500000,5000000,10000000")
args = ap.parse_args()
eng = Engine.get()
for n in [int(v) for v in args.sizes.split(",")]:
    ens = eng.alloc_ensemble(n, 1)
    gen = torch.Generator(device=eng.device).manual_seed(1)
    z = torch.randn(ens.ld, generator=gen, device=eng.device, dtype=torch.float64)
    ens.g.copy_(1.2 * z + 1.0)
    ens.e.copy_(0.5 * torch.randn(ens.ld, generator=gen, device=eng.device, dtype=torch.float64) ** 2 + 3.0)
    for label, sig_hi, do_beta in [("sigma+beta", 2.0 + 0.4, True), ("sigma", 2.4, False), ("beta", None, True),
                                   ("ess only", None, False)]:
        out = eng.adapt(ens, 0, 2.0, sig_hi, 2.0, 0.3, do_beta, 0.5 * n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            eng.adapt(ens, 0, 2.0, sig_hi, 2.0, 0.3, do_beta, 0.5 * n)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.reps
        evals = int(out[7])
        print(json.dumps({"n": n, "what": label, "us_per_call": round(us, 1), "combinations": evals,
                          "us_per_combination": round(us / evals, 2), "sigma_evals": int(out[5]),
                          "beta_evals": int(out[6]), "ctas": os.environ.get("REE_ADAPT_CTAS", "default")}))
