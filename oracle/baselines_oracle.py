"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the secondary paths: crude Monte
Carlo (``solver.py:1203-1271``) and SIS with adaptive conditional sampling
(``solver.py:1111-1147`` -> ``sis/SIS_aCS.py:64-306``).

Pinned against the live reference through ``tests/golden/make_golden.py``.
Only tests/, smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np
from scipy import optimize, stats

from lsf_oracle import normal_sample


def cmc(lsf, d, num_runs, seed=None, batch=2**10):
    """Crude Monte Carlo in batches of 1024 with per-batch seeds.  solver.py:1225-1271.

    Returns (pf, cvar, g) -- cvar = variation(1[g<=0]) = sqrt((1-p)/p) over all samples
    (the reference recomputes it over the whole prefix after every batch, :1252-1253;
    only the last value survives)."""
    rest = num_runs % batch
    nb = num_runs // batch + (1 if rest else 0)
    g = np.full(num_runs, np.nan)
    cost = 0
    for i in range(nb):
        bs = rest if (i == nb - 1 and rest) else batch
        bseed = None if seed is None else seed + i * num_runs  # :1242
        X = normal_sample(bs, d, bseed).reshape(bs, d)
        g[cost:cost + bs] = lsf(X)
        cost += bs
    fail = g <= 0
    pf = np.sum(fail) / cost
    with np.errstate(all="ignore"):
        cvar = stats.variation(fail, nan_policy="omit")
    return pf, cvar, g


def sis_acs(lsf, u0, p=0.1, burn=0, tar_cov=1.0, seed=None, max_it=100):
    """Sequential importance sampling with aCS-MCMC in standard-normal space.

    SIS_aCS.py:64-306 with ``transform_lsf=False`` (NormalProblem, solver.py:1144) and a
    given level-0 population ``u0`` (N,d).  Uses the *global* NumPy RNG like the reference
    (SIS_aCS.py:77, 194, 232, 248).  Returns (Pr, n_levels, last population, COV_Sl, n_evals)."""
    if seed is not None:
        np.random.seed(seed)
    N, dim = u0.shape
    if (N * p != np.fix(N * p)) or (1 / p != np.fix(1 / p)):
        raise RuntimeError("N*p and 1/p must be positive integers. Adjust N and p accordingly")
    n_evals = 0
    nchain = int(N * p)
    lenchain = int(N / nchain)
    adapchains = int(np.ceil(100 * nchain / N))
    lam = 0.6
    Sk = np.ones(max_it)
    sigmak = np.zeros(max_it + 1)
    uk = u0
    gk = lsf(uk)
    n_evals += N
    gmu = np.mean(gk)
    cdf = stats.norm.cdf
    m = 0
    cov_sl = np.nan
    for m in range(max_it):
        if m == 0:  # :151-165
            obj = lambda x: abs(np.std(cdf(-gk / x)) / np.mean(cdf(-gk / x)) - tar_cov)
            sigmak[m + 1] = optimize.fminbound(obj, 0, 10.0 * gmu)
            wk = cdf(-gk / sigmak[m + 1])
        else:  # :166-183
            den = cdf(-gk / sigmak[m])
            obj = lambda x: abs(np.std(cdf(-gk / x) / den) / np.mean(cdf(-gk / x) / den) - tar_cov)
            sigmak[m + 1] = optimize.fminbound(obj, 0, sigmak[m])
            wk = cdf(-gk / sigmak[m + 1]) / den
        Sk[m] = np.mean(wk)
        if Sk[m] == 0:
            break
        wnork = wk / Sk[m] / N
        ind = np.random.choice(range(N), nchain, True, wnork)  # :194
        gk0, uk0 = gk[ind], uk[ind, :]
        sigmaf = 1
        sigmafk = min(lam * sigmaf, 1)
        rhok = np.sqrt(1 - sigmafk**2)
        counta = 0
        count = 0
        alphak = np.zeros(nchain)
        gk = np.zeros(N)
        uk = np.zeros((N, dim))
        for k in range(nchain):  # :220-254
            u_cur, g_cur = uk0[k, :], gk0[k]
            for i in range(lenchain + burn):
                if i == burn:
                    count -= burn
                ucand = np.random.normal(loc=rhok * u_cur, scale=sigmafk)
                gcand = lsf(np.array(ucand, ndmin=2))
                n_evals += 1
                alpha = min(1, cdf(-gcand / sigmak[m + 1]) / cdf(-g_cur / sigmak[m + 1]))
                alphak[k] += alpha / (lenchain + burn)
                if stats.uniform.rvs() <= alpha:
                    uk[count, :] = ucand
                    gk[count] = gcand
                    u_cur, g_cur = ucand, gcand
                else:
                    uk[count, :] = u_cur
                    gk[count] = g_cur
                count += 1
            if (k + 1) % adapchains == 0:  # :257-268
                alpha_mu = np.mean(alphak[k - adapchains + 1:k + 1])
                counta += 1
                lam = min(np.exp(np.log(lam) + counta ** (-0.5) * (alpha_mu - 0.44)), 1)
                sigmafk = min(lam * sigmaf, 1)
                rhok = np.sqrt(1 - sigmafk**2)
        if sigmak[m + 1] == 0:
            cov_sl = np.nan
        else:
            r = (gk < 0) / cdf(-gk / sigmak[m + 1])
            cov_sl = np.std(r) / np.mean(r)
        if cov_sl < tar_cov:
            break
    n_levels = m + 1
    pr = np.mean((gk < 0) / cdf(-gk / sigmak[m + 1])) * np.prod(Sk)  # :294-300
    return pr, n_levels, uk, cov_sl, n_evals
