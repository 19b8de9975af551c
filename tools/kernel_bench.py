"""Times individual C-ABI entry points on synthetic data (CUDA events, warm-up, best/mean of R).
Usage: python tools/kernel_bench.py [--n 2000000] [--d 100] [--reps 5]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rareeventestimation_b200 import _lib  # noqa: E402
from rareeventestimation_b200.engine import DeviceLsf, Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2_000_000)
ap.add_argument("--d", type=int, default=100)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
n, d = args.n, args.d
eng = Engine.get()
X = eng.philox_normal(n, d, 1, 1)
Y = eng.alloc_ensemble(n, d)
lsf = DeviceLsf(_lib.LSF_AFFINE, d, None, 3.5)
eng.lsf_eval(lsf, X, out=X.g)
eng.energy(X, out=X.e)
slen = d + d * d + 2
sums, sums_w = eng.empty(slen), eng.empty(slen)
mean, cov, L, Linv, stats = eng.empty(d), eng.empty(d * d), eng.empty(d * d), eng.empty(d * d), eng.empty(8)
ess4, is4 = eng.empty(4), eng.empty(4)
m_noise = eng.zeros(d)
F = eng.empty(d * d)


def timeit(name, fn, bytes_alg=None, flops=None):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    out = {"kernel": name, "ms": round(ms, 4), "n": n, "d": d}
    if bytes_alg:
        out["GB/s"] = round(bytes_alg / ms / 1e6, 1)
    if flops:
        out["TFLOP/s"] = round(flops / ms / 1e9, 2)
    print(json.dumps(out), flush=True)


gram_flops = n * (d + 1) * (d + 2)  # symmetric half, 2 flops per FMA
tri_flops = n * d * (d + 1)
timeit("ree_ensemble_sums", lambda: eng.ensemble_sums(X, None, sums), n * 8 * d, gram_flops)
eng.moments_chol(sums, None, d, 1, False, mean, cov, L, Linv, stats)
timeit("ree_moments_chol", lambda: eng.moments_chol(sums, None, d, 1, False, mean, cov, L, Linv, stats))
timeit("ree_chol_scaled", lambda: eng.chol_scaled(cov, d, 0.5, F, stats))
timeit("ree_reduce_ess", lambda: eng.reduce_ess(X, 2.0, 1.0, 0, out=ess4), n * 16)
timeit("ree_reduce_cvar", lambda: eng.reduce_cvar(X, 2.5, 2.0, 0, out=ess4), n * 8)
eng.reduce_ess(X, 2.0, 1.0, 0, out=ess4)
timeit("ree_cbree_maha", lambda: eng.cbree_maha(X, mean, Linv, stats[0:1], 2.0, 1.0, 0, ess4[0:1], sums_w, is4),
       n * (8 * d + 16), gram_flops + tri_flops)
timeit("ree_cbree_step(philox)", lambda: eng.cbree_step(X, Y, 0.6, m_noise, F, lsf, None, sums, seed=7, draw=3),
       n * (16 * d + 16), gram_flops + tri_flops)
if eng.path(d) == 2:
    a_vec = eng.empty(X.ld)
    sums_uw = eng.empty(2 * slen)
    eng.reduce_ess(X, 2.0, 1.0, 0, out=ess4, a_out=a_vec)
    timeit("split: ree_cbree_step(philox, transform only)",
           lambda: eng.cbree_step(X, Y, 0.6, m_noise, F, lsf, None, None, seed=7, draw=3), n * (16 * d + 16), tri_flops)
    timeit("split: ree_cbree_moments(both Grams)",
           lambda: eng.cbree_moments(X, a_vec, ess4[0:1], None, sums_uw), n * (8 * d + 16), 2 * gram_flops)
    timeit("split: ree_cbree_moments(unweighted)",
           lambda: eng.cbree_moments(X, None, None, None, sums_uw, weighted=False), n * (8 * d + 8), gram_flops)
    timeit("split: ree_cbree_maha(transform only)",
           lambda: eng.cbree_maha(X, mean, Linv, stats[0:1], 2.0, 1.0, 0, None, None, is4), n * (8 * d + 16), tri_flops)
Z = eng.philox_normal(n, d, 2, 2)
timeit("ree_cbree_step(inject)", lambda: eng.cbree_step(X, Y, 0.6, m_noise, F, lsf, None, sums, z=Z),
       n * (24 * d + 16), gram_flops + 2 * n * d * d)
timeit("ree_philox_normal", lambda: eng.philox_normal(n, d, 3, 3, out=Z), n * 8 * d)
timeit("ree_lsf_eval(affine)", lambda: eng.lsf_eval(lsf, X, out=X.g), n * (8 * d + 8))
timeit("ree_energy", lambda: eng.energy(X, out=X.e), n * (8 * d + 8))
# von Mises-Fisher-Nakagami model (resampling branch)
from rareeventestimation_b200.mixture import VMFNMixture, _house  # noqa: E402

Xs = eng.philox_normal(n, d, 5, 5)
Xs.X.add_(1.0)  # a population away from the origin: a well-defined mean direction
model = VMFNMixture(1)
timeit("ree_vmfn_fit_sums", lambda: eng.vmfn_fit_sums(Xs), n * 16 * d)
model.fit_device(eng, Xs)
hv, hb = _house(model.mu[:, 0])
hvd = eng.to_device(hv)
kap, mm, om = model._scalars()
timeit("ree_vmfn_sample(philox)", lambda: eng.vmfn_sample(Y, hvd, hb, kap, mm, om, 3, 4), n * 8 * d)
eng.lsf_eval(lsf, Y, out=Y.g)
eng.energy(Y, out=Y.e)
mud = eng.to_device(np.ascontiguousarray(model.mu[:, 0]))
timeit("ree_vmfn_logpdf(+IS statistics)", lambda: eng.vmfn_logpdf(Y, mud, kap, mm, om, model._c0(), is4=is4),
       n * (8 * d + 16))
timeit("ree_vector_range", lambda: eng.vector_range(X.g, n, d), n * 8)
rng2 = eng.vector_range(X.e, n, d)
timeit("ree_reduce_scaled(known shift)", lambda: eng.reduce_scaled(X.e, n, d, -0.3, out=ess4, range2=rng2), n * 8)
timeit("ree_reduce_scaled(running shift)", lambda: eng.reduce_scaled(X.e, n, d, -0.3, out=ess4), n * 8)
