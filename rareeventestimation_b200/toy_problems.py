"""The reference's toy problems (``problem/toy_problems.py``) with device limit-state functions."""
from __future__ import annotations

import numpy as np
from scipy.stats import gamma, norm

from .lsf import AffineLSF, ConvexLSF, FujitaRackwitzLSF, OscillatorLSF
from .problem import NormalProblem

sample_size = 1  # replaced through set_sample by every caller, as in the reference

lsf_convex = ConvexLSF()
prob_convex = NormalProblem(lsf_convex, 2, sample_size, prob_fail_true=4.21e-3, name="Convex Problem",
                            mpp=5 / (2 * np.sqrt(2)) * np.ones(2))
prob_convex.matlab_str = "@(x) 0.1*(x(:,1)-x(:,2)).^2 - 1/sqrt(2)*(x(:,1)+x(:,2))+2.5"

beta = 3.5
lsf_lin = AffineLSF(beta)


def make_linear_problem(d, beta=3.5):
    """g(x) = -sum(x)/sqrt(d) + beta, P_f = Phi(-beta)   (toy_problems.py:35-48)."""
    p = NormalProblem(AffineLSF(beta), d, 1, prob_fail_true=norm.cdf(-beta), name=f"Linear Problem (d={d})",
                      mpp=beta / np.sqrt(d) * np.ones(d))
    p.matlab_str = f"@(x) -sum(x)/sqrt(length(x)) + {beta}"
    return p


def make_fujita_rackwitz(d, pf=1e-04):
    """toy_problems.py:54-66."""
    c = gamma.ppf(pf, d)
    prob = NormalProblem(FujitaRackwitzLSF(c), d, 1, gamma.cdf(c, d), name=f"Fujita Rackwitz Problem (d={d})",
                         mpp=-norm.ppf(np.exp(-c / d)) * np.ones(d))
    prob.matlab_str = f"@(x) -sum(log(normcdf(-x))) - {c}"
    return prob


lsf_nonlin_osc = OscillatorLSF()
prob_nonlin_osc = NormalProblem(lsf_nonlin_osc, 6, 1, prob_fail_true=6.43 * 1e-6,
                                name="Nonlinear Oscillator Problem")

problems_lowdim = [prob_convex, make_linear_problem(2), make_fujita_rackwitz(2)]
problems_highdim = [make_linear_problem(d) for d in [10, 20, 30, 40, 50, 75, 100, 150, 200]]
problems_highdim.extend([make_fujita_rackwitz(d) for d in [10, 20, 30, 40, 50, 75, 100, 150, 200]])
