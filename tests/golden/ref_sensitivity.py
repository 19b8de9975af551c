"""Sensitivity of the REFERENCE's own diffusion trajectory (N=300, d=150, seeds 2/2) to a perturbation of the initial
sample in the last bits: runs the live reference (build container only) twice and compares the caches.
Usage: python tests/golden/ref_sensitivity.py [relative perturbation, default 1e-14]"""
import os
import sys
from copy import deepcopy

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # tests/golden -> repo root
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_loader  # noqa: E402

ref = ref_loader.load_reference()
dm = ref_loader.load_reference_diffusion()
eps = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-14


def run(perturb):
    prob = deepcopy(dm.diffusion_problem)
    prob.set_sample(300, seed=2)
    if perturb:
        rng = np.random.default_rng(99)
        prob.sample = prob.sample * (1 + eps * rng.standard_normal(prob.sample.shape))
    solver = ref.solver.CBREE(seed=2, save_history=True, return_caches=True, num_steps=8)
    return solver.solve(prob).other["cache_list"]


a, b = run(False), run(True)
for i, (ca, cb) in enumerate(zip(a, b)):
    dx = np.abs(ca.ensemble - cb.ensemble).max()
    sa = np.linalg.svd((1 - ca.t_step**2) * (1 + ca.beta) * ca.weighted_cov)
    sb = np.linalg.svd((1 - cb.t_step**2) * (1 + cb.beta) * cb.weighted_cov)
    flips = int((np.sum(sa[0] * sb[0], axis=0) < 0).sum())
    print(f"cache {i}: |dX| max {dx:.3e}  beta {ca.beta:.10f} / {cb.beta:.10f}  t {ca.t_step:.3e}  "
          f"flipped U columns of the next step's factor: {flips}")
