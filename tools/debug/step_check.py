"""Engine step 6 -> 7 of the golden diffusion run against the NumPy formula on the engine's own inputs."""
import json
import os
import sys
from copy import deepcopy

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rareeventestimation_b200 as ree  # noqa: E402

t = np.load(os.path.join(ROOT, "tests", "golden", "cbree_diffusion_d150.npz"))
cfg = json.loads(str(t["config"]))
p = deepcopy(ree.diffusion_problem).set_sample(cfg["n"], seed=cfg["sample_seed"])
solver = ree.CBREE(seed=cfg["solver_seed"], save_history=True, return_caches=True, noise="numpy", quiet=True, **cfg["opts"])


class Rng:
    def __init__(self, g):
        self.g, self.draws = g, []

    def standard_normal(self, shape):
        z = self.g.standard_normal(shape)
        self.draws.append(z)
        return z

    def __getattr__(self, k):
        return getattr(self.g, k)


solver.rng = Rng(solver.rng)
sol = solver.solve(p)
caches = sol.other["cache_list"]
print("draws", len(solver.rng.draws), "caches", len(caches), [c.iteration for c in caches])
for i in range(1, len(caches)):
    a, b = caches[i - 1], caches[i]
    if b.iteration != a.iteration + 1:
        continue
    Z = solver.rng.draws[b.iteration - 1 + (len(solver.rng.draws) - caches[-1].iteration)]
    tt = float(a.t_step)
    X = np.asarray(a.ensemble)
    c_noise = (1 - tt**2) * (1 + a.beta) * np.asarray(a.weighted_cov)
    u, s, _ = np.linalg.svd(c_noise)
    F = u * np.sqrt(s)
    want = tt * X + ((1 - tt) * np.asarray(a.weighted_mean) + Z @ F.T)
    got = np.asarray(b.ensemble)
    print(f"step {a.iteration}->{b.iteration}: t={tt:.3e} beta={a.beta:.6f} |got-want| max {np.abs(got - want).max():.3e}  |want| max {np.abs(want).max():.3f}")
