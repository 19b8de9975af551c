"""Logs MINPACK hybr's exit status of every beta update of the golden diffusion run (sensitivity of the trace test)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rareeventestimation_b200 as ree  # noqa: E402
from rareeventestimation_b200 import solver as S  # noqa: E402
from copy import deepcopy  # noqa: E402

t = np.load(os.path.join(ROOT, "tests", "golden", "cbree_diffusion_d150.npz"))
cfg = json.loads(str(t["config"]))
orig = S.root
log = []


def root_logged(fun, x0, *a, **k):
    vals = []

    def f(x):
        v = fun(x)
        vals.append((float(np.asarray(x).reshape(-1)[0]), float(v)))
        return v
    if len(log) == 8:
        cells = {n: c.cell_contents for n, c in zip(fun.__code__.co_freevars, fun.__closure__)}
        cache, ell = cells["cache"], cells["ell"]
        dev = cache._dev
        np.savez(os.path.join(ROOT, "gpurun_out", "eng_upd8.npz"), g=dev.g[:dev.n].cpu().numpy(), e=dev.e[:dev.n].cpu().numpy(),
                 sigma=cache.sigma, ell=ell[:dev.n].cpu().numpy(), X=cells["self"]._eng.download(dev))
        print("saved", cache.sigma)
    sol = orig(f, x0, *a, **k)
    log.append((float(x0), sol.x.item(), sol.success, sol.status, sol.nfev, float(sol.fun[0]) if hasattr(sol.fun, "__len__") else float(sol.fun)))
    if len(log) in (8, 9):
        for x, v in vals:
            print(f"      beta={x!r} f={v!r}")
    return sol


S.root = root_logged
p = deepcopy(ree.diffusion_problem).set_sample(cfg["n"], seed=cfg["sample_seed"])
solver = ree.CBREE(seed=cfg["solver_seed"], save_history=True, return_caches=True, noise="numpy", quiet=True, **cfg["opts"])
sol = solver.solve(p)
for i, l in enumerate(log):
    print(i, l)
print("golden beta", t["beta"])

# second pass: cross-check every ESS evaluation against NumPy on the downloaded log-target vector
print("---- cross-check")
S.root = orig
_ess0 = S.CBREE._ess
worst = [0.0]


def ess_checked(self, cache, beta, out=None, ell=None):
    v = _ess0(self, cache, beta, out, ell)
    if ell is not None:
        n = cache._dev.n
        l = ell[:n].cpu().numpy()
        a = l * beta
        a = a[np.isfinite(a)]
        w = np.exp(a - a.max())
        ref = w.sum() ** 2 / (w * w).sum()
        err = abs(v - ref) / ref
        if err > worst[0]:
            worst[0] = err
            print(f"   it={cache.iteration} beta={beta!r} gpu={v!r} numpy={ref!r} rel={err:.3e} n={n} finite={a.size} min_a={a.min():.3f}")
    return v


S.CBREE._ess = ess_checked
p = deepcopy(ree.diffusion_problem).set_sample(cfg["n"], seed=cfg["sample_seed"])
solver = ree.CBREE(seed=cfg["solver_seed"], save_history=True, return_caches=True, noise="numpy", quiet=True, **cfg["opts"])
sol = solver.solve(p)
print("worst", worst, cfg)
caches = sol.other["cache_list"]
for key in ["sigma", "beta", "t_step", "sfp", "ess", "cvar_is_weights"]:
    if key in t.files:
        got = np.array([getattr(c, key) for c in caches], dtype=np.float64)
        print(key, "\n  got   ", got, "\n  golden", t[key])
print(t.files)
