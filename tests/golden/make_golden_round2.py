"""Round-2 golden fixtures from the LIVE reference (build container only: needs /root/reference or oracle/_ref).

    python tests/golden/make_golden_round2.py [singular] [pf]

singular  near-singular Gaussian fits: ensembles whose covariance has lambda_min / lambda_max around SciPy's threshold
          1e6 * eps (scipy _PSD inside multivariate_normal.logpdf, solver.py:672, 838): how the reference's
          CBREE(t_step=0.5).solve ends for each -- the exception it raises, or Solution.msg -> singular_cases.npz
pf        distribution of the final estimate over 200 solver seeds on a fixed sample (the loop of do_multiple_solves,
          evaluation/convergence_analysis.py:34-148: set_options({"seed": i, "rng": default_rng(i)}), solve) for the
          README problem d = 4 N = 2000, the oscillator N = 5000 and the linear problem d = 50 with vMFNM resampling
          -> pf_distribution_*.npz
"""
import json
import os
import sys
import warnings
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_loader  # noqa: E402

ref = ref_loader.load_reference()
warnings.simplefilter("ignore")


def singular_cases():
    rng = np.random.default_rng(2024)
    n, d = 500, 3
    Z = rng.standard_normal((n, d))
    Q, _ = np.linalg.qr(rng.standard_normal((d, d)))  # the thin direction is not axis aligned
    scales2 = [1e-6, 1e-8, 2e-9, 9e-10, 5e-10, 3.2e-10, 1.6e-10, 1.0e-10, 3e-11, 1e-12, 1e-14, 0.0]
    out = {"scales2": np.array(scales2), "n": n, "d": d}
    outcomes = []
    for k, s2 in enumerate(scales2):
        X = (Z * np.array([1.0, 1.0, np.sqrt(s2)])) @ Q.T
        prob = ref.toy.make_linear_problem(d)
        prob.set_sample(X)
        lam = np.linalg.eigvalsh(np.cov(X, rowvar=False))
        try:
            sol = ref.solver.CBREE(seed=1, t_step=0.5, num_steps=8).solve(prob)
            res = {"raises": None, "msg": sol.msg, "num_steps": int(sol.num_steps),
                   "pf": float(sol.prob_fail_hist[-1])}
        except Exception as e:  # the final importance sampling re-fits every retained cache (solver.py:473-474)
            res = {"raises": type(e).__name__, "msg": str(e)}
        res["eig_ratio_x0"] = float(lam[0] / lam[-1])
        outcomes.append(res)
        out[f"X_{k}"] = X
        print(f"  s2={s2:8.1e}  lambda_min/lambda_max={res['eig_ratio_x0']:.3e}  ->  {res}")
    out["outcomes"] = json.dumps(outcomes)
    np.savez_compressed(os.path.join(HERE, "singular_cases.npz"), **out)


def pf_distribution(name, make, n, sample_seed, runs=200, **opts):
    prob = make()
    prob.set_sample(n, seed=sample_seed)
    base = ref.solver.CBREE(**opts)
    est, steps, msgs, costs = [], [], [], []
    for i in range(runs):
        solver = base.set_options({"seed": i, "rng": np.random.default_rng(i)}, in_situ=False)
        sol = solver.solve(prob)
        est.append(sol.prob_fail_hist[-1])
        steps.append(sol.num_steps)
        costs.append(sol.costs)
        msgs.append(sol.msg)
    est = np.array(est)
    truth = prob.prob_fail_true
    print(f"  {name}: mean {est.mean():.4e} (truth {truth:.4e}), rel. RMSE {np.sqrt(np.mean((est - truth) ** 2)) / truth:.4f}, "
          f"success {np.mean([m == 'Success' for m in msgs]):.2f}, steps {np.mean(steps):.1f}")
    np.savez_compressed(os.path.join(HERE, f"pf_distribution_{name}.npz"), estimates=est, steps=np.array(steps),
                        costs=np.array(costs), msgs=np.array(msgs), truth=truth,
                        config=json.dumps(dict(n=n, sample_seed=sample_seed, runs=runs, opts=opts)))


if __name__ == "__main__":
    what = sys.argv[1:] or ["singular", "pf"]
    if "singular" in what:
        singular_cases()
    if "pf" in what:
        pf_distribution("readme_d4", lambda: ref.toy.make_linear_problem(4), 2000, 1)
        pf_distribution("osc_d6", lambda: deepcopy(ref.toy.prob_nonlin_osc), 5000, 1)
        pf_distribution("linear_d50_vmfnm", lambda: ref.toy.make_linear_problem(50), 5000, 5000,
                        resample=True, mixture_model="vMFNM")
