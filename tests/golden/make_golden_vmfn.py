"""Golden fixtures of the vMFNM resampling branch from the LIVE reference (build container only):

    python tests/golden/make_golden_vmfn.py

* vmfn_vectors.npz: VMFNMixture(1).fit / logpdf / sample (mixturemodel.py:153-239, era/EMvMFNM.py) on seeded inputs,
  d in {2, 6, 50}, uniform and non-uniform weights;
* cbree_vmfnm_*.npz: CBREE(resample=True, mixture_model="vMFNM") traces (solver.py:654-663, 828-836), written by the
  same trace_cbree() as the other runs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (loads the reference; its __main__ block does not run)

ref = mg.ref
mm = __import__("importlib").import_module("rareeventestimation.mixturemodel")


def vmfn_vectors():
    out = {}
    cases = [(2, 400, 11), (6, 500, 12), (50, 700, 13)]
    out["cases"] = np.array(cases)
    for d, n, seed in cases:
        rng = np.random.default_rng(seed)
        centre = rng.standard_normal(d) * 0.7 + 1.0
        X = centre + rng.standard_normal((n, d)) * np.linspace(0.5, 1.5, d)
        w = rng.random(n)
        w /= w.sum()
        Y = rng.standard_normal((64, d)) * 1.3 + 0.5 * centre
        for tag, weights in (("u", None), ("w", w)):
            model = mm.VMFNMixture(1)
            np.random.seed(0)  # EMvMFNM's initialisation touches the global generator
            model.fit(X, weights)
            key = f"d{d}_{tag}_"
            out[key + "mu"], out[key + "kappa"], out[key + "m"] = model.mu, model.kappa, model.m
            out[key + "omega"], out[key + "alpha"] = model.omega, model.alpha
            out[key + "logpdf_X"] = model.logpdf(X[:64])
            out[key + "logpdf_Y"] = model.logpdf(Y)
            out[key + "sample"] = model.sample(200, rng=np.random.default_rng(seed + 100))
        out[f"d{d}_X"], out[f"d{d}_w"], out[f"d{d}_Y"] = X, w, Y
    np.savez_compressed(os.path.join(HERE, "vmfn_vectors.npz"), **out)
    print("vmfn_vectors done:", {k: v.shape for k, v in out.items() if k.endswith("kappa")})


if __name__ == "__main__":
    vmfn_vectors()
    # the reference's own use (test/test_cbree.py:13-25) at sizes it finishes in seconds
    mg.trace_cbree("cbree_vmfnm_linear_d10", "affine", 10, 800, 5, 1, mat_keep=2, resample=True, mixture_model="vMFNM",
                   num_steps=12)
    mg.trace_cbree("cbree_vmfnm_fr_d20", "fujita_rackwitz", 20, 600, 6, 2, mat_keep=2, resample=True,
                   mixture_model="vMFNM", num_steps=10)
    # importance sampling with the model but Gaussian dynamics (solver.py:828-836)
    mg.trace_cbree("cbree_vmfnm_is_only_d6", "affine", 6, 600, 7, 3, mat_keep=2, mixture_model="vMFNM", num_steps=10)
