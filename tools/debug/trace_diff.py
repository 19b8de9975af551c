"""Per-iteration differences between an engine solve and a golden trace.  Usage: trace_diff.py <golden name>"""
import json
import os
import sys
from copy import deepcopy

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rareeventestimation_b200 as ree  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cbree_diffusion_d150"
t = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
cfg = json.loads(str(t["config"]))
kind = cfg["kind"]
p = {"diffusion": lambda: deepcopy(ree.diffusion_problem), "affine": lambda: ree.make_linear_problem(cfg["d"]),
     "oscillator": lambda: deepcopy(ree.prob_nonlin_osc)}[kind]()
p.set_sample(cfg["n"], seed=cfg["sample_seed"])
solver = ree.CBREE(seed=cfg["solver_seed"], save_history=True, return_caches=True, noise="numpy", quiet=True, **cfg["opts"])
sol = solver.solve(p)
caches = sol.other["cache_list"]
for i, c in enumerate(caches):
    ens = np.asarray(c.ensemble)
    k = t["ens_head"].shape[1]
    dx = np.abs(ens[:k] - t["ens_head"][i]).max()
    dg = np.abs(np.asarray(c.lsf_evals)[:t["lsf_head"].shape[1]] - t["lsf_head"][i]).max()
    dm = np.abs(np.asarray(c.weighted_mean) - t["weighted_mean"][i]).max()
    de = np.abs(ens.mean(axis=0) - t["ens_mean"][i]).max()
    print(f"cache {i}: |dX head| {dx:.3e}  |dg head| {dg:.3e}  |d m_w| {dm:.3e}  |d ens_mean| {de:.3e}  beta {c.beta!r} vs {t['beta'][i]!r}  t {c.t_step:.3e}")
# sensitivity of the reference's noise factor U sqrt(S) (Generator.multivariate_normal(method="svd")) to the covariance
rng = np.random.default_rng(0)
for i in (1, 3, 5, 6):
    C = np.asarray(caches[i].weighted_cov)
    u, s, _ = np.linalg.svd(C)
    F = u * np.sqrt(s)
    E = rng.standard_normal(C.shape)
    E = (E + E.T) * 0.5e-9 * np.abs(C).max()
    u2, s2, _ = np.linalg.svd(C + E)
    F2 = u2 * np.sqrt(s2)
    # column signs: a flipped column of U is a different (equally distributed) sample Z F^T
    sg = np.sign(np.sum(F * F2, axis=0))
    print(f"   flipped columns of U sqrt(S) under the 1e-9 perturbation: {(sg < 0).sum()} of {sg.size}")
    gaps = np.abs(np.diff(s)) / s[:-1]
    print(f"cache {i}: s max {s[0]:.3e} min {s[-1]:.3e}  #s<1e-12*max {(s < 1e-12 * s[0]).sum()}  min rel gap {gaps.min():.2e}  "
          f"|dF|/|F| after 1e-9 perturbation {np.linalg.norm(F - F2 * sg) / np.linalg.norm(F):.2e}")
