"""Generate the golden fixtures under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture is an ``.npz`` holding the configuration (so the oracle / CUDA
tests can rebuild the inputs from seeds) and the reference's per-iteration
results.  The reference itself cannot travel to the GPU box; these files do.
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

HEAD = 32  # leading entries of N-vectors kept verbatim


def make_problem(kind, d):
    if kind == "affine":
        return ref.toy.make_linear_problem(d)
    if kind == "convex":
        return deepcopy(ref.toy.prob_convex)
    if kind == "fujita_rackwitz":
        return ref.toy.make_fujita_rackwitz(d)
    if kind == "oscillator":
        return deepcopy(ref.toy.prob_nonlin_osc)
    if kind == "diffusion":
        return deepcopy(ref_loader.load_reference_diffusion().diffusion_problem)
    raise ValueError(kind)


def trace_cbree(name, kind, d, n, sample_seed, solver_seed, mat_keep=1000, **opts):
    prob = make_problem(kind, d)
    prob.set_sample(n, seed=sample_seed)
    solver = ref.solver.CBREE(seed=solver_seed, save_history=True, return_caches=True, return_other=True, **opts)
    sol = solver.solve(prob)
    caches = sol.other["cache_list"]
    out = {
        "config": json.dumps(
            dict(kind=kind, d=d, n=n, sample_seed=sample_seed, solver_seed=solver_seed, opts=opts)
        ),
        "msg": sol.msg,
        "costs": sol.costs,
        "num_steps": sol.num_steps,
        "prob_fail_hist": np.asarray(sol.prob_fail_hist),
        "temp_hist": np.asarray(sol.temp_hist),
        "avg_estimates": np.array(
            [
                sol.other["Average Estimate"],
                sol.other["Root Weighted Average Estimate"],
                sol.other["VAR Weighted Average Estimate"],
            ]
        ),
        "x0_head": prob.sample[:HEAD].copy(),
        "x0_sum": prob.sample.sum(axis=0),
    }
    scal = ["sigma", "beta", "t_step", "cvar_is_weights", "sfp", "ess", "estimate", "iteration",
            "converged", "slope_cvar", "sfp_slope", "num_lsf_evals"]
    for s in scal:
        out[s] = np.array([getattr(c, s) for c in caches], dtype=np.float64)
    out["weighted_mean"] = np.array([c.weighted_mean for c in caches])
    out["weighted_cov"] = np.array([c.weighted_cov for c in caches[:mat_keep]])
    out["ens_mean"] = np.array([c.ensemble.mean(axis=0) for c in caches])
    out["ens_cov"] = np.array([np.atleast_2d(np.cov(c.ensemble, rowvar=False, ddof=1)) for c in caches[:mat_keep]])
    out["ens_head"] = np.array([c.ensemble[:HEAD if d <= 10 else 4] for c in caches])
    out["lsf_head"] = np.array([c.lsf_evals[:HEAD] for c in caches])
    out["lsf_sum"] = np.array([c.lsf_evals.sum() for c in caches])
    out["efun_head"] = np.array([c.e_fun_evals[:HEAD] for c in caches])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: steps={sol.num_steps} pf={sol.prob_fail_hist[-1]:.6e} msg={sol.msg!r} costs={sol.costs}")


def unit_vectors():
    """Known-answer vectors for the N-vector helpers (utilities.py) on fixed inputs."""
    rng = np.random.default_rng(12345)
    g = rng.standard_normal(257) * 2 + 1
    e = 0.5 * rng.standard_normal((257, 5)).__pow__(2).sum(1) + 2.5 * np.log(2 * np.pi)
    out = {"g": g, "e": e}
    s = ref.solver.CBREE(seed=0)
    ltf = s._CBREE__log_tgt_fun
    for kind in ["algebraic", "tanh", "arctan", "erf", "relu", "sigmoid"]:
        for sig in [0.0, 0.7, 35.0]:
            out[f"logtgt_{kind}_{sig}"] = ltf(g, e, sig, method=kind)
    lt = ltf(g, e, 3.0, method="algebraic")
    for beta in [0.25, 1.0, 7.5]:
        w = ref.utilities.my_softmax(lt * beta)
        out[f"softmax_{beta}"] = w
    out["softmax_degenerate"] = ref.utilities.my_softmax(np.array([-1e6, -2e6, -np.inf, -3e6]))
    out["softmax_nonfinite"] = ref.utilities.my_softmax(np.array([0.5, np.nan, -0.5, np.inf, 0.1]))
    out["cvar_plain"] = ref.utilities.my_log_cvar(lt)
    out["cvar_mult"] = ref.utilities.my_log_cvar(lt, multiplier=(g <= 0))
    out["cvar_zero"] = ref.utilities.my_log_cvar(lt, multiplier=np.zeros_like(g))
    out["is_est"] = ref.utilities.importance_sampling(-e, lt, g, logpdf=True)
    out["slope"] = ref.utilities.get_slope(np.array([3.0, 2.5, 2.7, 1.0, 0.5]))
    out["slope_nan"] = ref.utilities.get_slope(np.array([3.0, np.nan, 2.7]))
    X = rng.standard_normal((64, 6))
    out["X"] = X
    out["efun"] = ref.toy.prob_nonlin_osc.e_fun(X)
    out["lsf_osc"] = ref.toy.lsf_nonlin_osc(X)
    Xo = X.copy()
    Xo[:4, 1] = -12.0  # force omega^2 < 0 (complex branch)
    Xo[:4, 2] = -30.0
    out["X_osc_neg"] = Xo
    out["lsf_osc_neg"] = ref.toy.lsf_nonlin_osc(Xo)
    out["lsf_convex"] = ref.toy.lsf_convex(X[:, :2])
    out["lsf_affine"] = ref.toy.make_linear_problem(6).lsf(X)
    fr = ref.toy.make_fujita_rackwitz(6)
    out["lsf_fr"] = fr.lsf(X)
    out["fr_truth"] = fr.prob_fail_true
    p = ref.toy.make_linear_problem(3)
    p.set_sample(50, seed=7)
    out["sample_50x3_seed7"] = p.sample
    np.savez_compressed(os.path.join(HERE, "unit_vectors.npz"), **out)
    print("unit_vectors done")


def diffusion_vectors():
    dm = ref_loader.load_reference_diffusion()
    rng = np.random.default_rng(2024)
    theta = rng.standard_normal((24, 150))
    theta[0] = 0.0
    theta[1] *= 3.0
    g = np.array([dm.g(t) for t in theta])
    np.savez_compressed(
        os.path.join(HERE, "diffusion_vectors.npz"),
        theta=theta,
        g=g,
        mode_table_head=dm.function_values[:9, :],
        mode_table_tail=dm.function_values[-9:, :],
        mode_table_colsum=dm.function_values.sum(axis=0),
        mu=dm.mu,
        sigma=dm.sigma,
    )
    print("diffusion_vectors done; g(0) =", g[0])


def baselines():
    # CMC on the convex problem (test/test_cmc.py:4-8, shortened)
    prob = make_problem("convex", 2)
    sol = ref.solver.CMC(20000, seed=1, verbose=False).solve(prob)
    cmc = dict(pf=sol.prob_fail_hist[-1], cvar=sol.other["cvar"], costs=sol.costs,
               lsf_head=sol.lsf_eval_hist[0, :HEAD], lsf_tail=sol.lsf_eval_hist[0, -HEAD:])
    # SIS aCS on linear d=2 (test/test_sis.py:5-14, shortened)
    prob = make_problem("affine", 2)
    prob.set_sample(1000, seed=1000)
    sol = ref.solver.SIS(seed=1, mixture_model="aCS", return_other=True).solve(prob)
    sis = dict(pf=sol.prob_fail_hist[-1], levels=sol.num_steps, costs=sol.costs, cov_sl=sol.other["COV_Sl"],
               last_head=sol.ensemble_hist[0, :HEAD])
    np.savez_compressed(os.path.join(HERE, "baselines.npz"),
                        **{"cmc_" + k: v for k, v in cmc.items()}, **{"sis_" + k: v for k, v in sis.items()})
    print("baselines: cmc pf", cmc["pf"], "sis pf", sis["pf"], "levels", sis["levels"], "costs", sis["costs"])


if __name__ == "__main__":
    unit_vectors()
    diffusion_vectors()
    baselines()
    # README example (BASELINE.json configs[0])
    trace_cbree("cbree_readme_d4", "affine", 4, 2000, 1, 1)
    # reference tests: test/test_cbree.py:5-10
    trace_cbree("cbree_linear_d2", "affine", 2, 5000, 5000, 1, divergence_check=False)
    trace_cbree("cbree_convex_d2", "convex", 2, 5000, 5000, 1, divergence_check=False)
    trace_cbree("cbree_fr_d2", "fujita_rackwitz", 2, 5000, 5000, 1, divergence_check=False)
    # test/test_nonlinear_oscillator_prob.py:4-9 (seeded here)
    trace_cbree("cbree_osc_d6", "oscillator", 6, 5000, 1, 1)
    # headline shape at oracle size
    trace_cbree("cbree_linear_d100", "affine", 100, 2000, 1, 1, mat_keep=4)
    # option coverage
    trace_cbree("cbree_d4_tanh_sigma5", "affine", 4, 1000, 3, 3, tgt_fun="tanh", sigma_adaptivity=5.0, num_steps=10)
    trace_cbree("cbree_d4_sfp", "affine", 4, 1000, 3, 3, sigma_adaptivity="sfp", num_steps=15)
    trace_cbree("cbree_d4_arctan_fixed", "affine", 4, 1000, 4, 4, tgt_fun="arctan", t_step=0.5, num_steps=12)
    trace_cbree("cbree_d4_erf_beta2", "affine", 4, 1000, 5, 5, tgt_fun="erf", beta_adaptivity=2.0, num_steps=15)
    trace_cbree("cbree_d4_relu", "affine", 4, 1000, 6, 6, tgt_fun="relu", num_steps=15)
    trace_cbree("cbree_d4_sigmoid", "affine", 4, 1000, 7, 7, tgt_fun="sigmoid", num_steps=15,
                observation_window=3)
    # PDE LSF through CBREE at a size the reference finishes in seconds
    trace_cbree("cbree_diffusion_d150", "diffusion", 150, 300, 2, 2, mat_keep=2, num_steps=8)
