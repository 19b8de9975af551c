"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the limit-state functions and the
initial N(0,I) sample of the reference.

Restates ``src/rareeventestimation/problem/toy_problems.py``,
``problem/diffusion.py`` and ``era/ERANataf.py:372-386`` in NumPy.  Pinned against
the live reference by ``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``.
Only tests/, smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import stats


# --------------------------------------------------------------------------- #
# initial sample: NormalProblem.sample_gen -> ERANataf.random (ERANataf.py:372-386)
# --------------------------------------------------------------------------- #
def normal_sample(n: int, d: int, seed=None) -> np.ndarray:
    """(n,d) sample as the reference draws it: dimension-major N(0,1) block, then
    the per-marginal ppf(cdf(.)) round trip of the Nataf transform."""
    rng = np.random.default_rng(seed)
    U = rng.standard_normal((d, int(n)))
    out = np.empty((int(n), d))
    for j in range(d):
        out[:, j] = stats.norm.ppf(stats.norm.cdf(U[j]))
    return np.squeeze(out) if n == 1 else out


# --------------------------------------------------------------------------- #
# toy limit-state functions (toy_problems.py)
# --------------------------------------------------------------------------- #
def lsf_affine(x, beta=3.5):
    """toy_problems.py:36-37 / README.md:31-49."""
    return -np.sum(x, axis=-1) / np.sqrt(x.shape[-1]) + beta


def lsf_convex(x):
    """toy_problems.py:12-13."""
    return 0.1 * (x[..., 0] - x[..., 1]) ** 2 - 1 / np.sqrt(2) * np.sum(x, axis=-1) + 2.5


def fujita_rackwitz_offset(d, pf=1e-4):
    """toy_problems.py:55."""
    return stats.gamma.ppf(pf, d)


def lsf_fujita_rackwitz(x, c):
    """toy_problems.py:57-58."""
    return -np.sum(np.log(stats.norm.cdf(-x)), axis=-1) - c


def lsf_oscillator(x):
    """toy_problems.py:72-80 (complex sqrt so that omega^2 < 0 is handled)."""
    m = 1 + 0.05 * x[..., 0]
    c1 = 1 + 0.1 * x[..., 1]
    c2 = 0.1 + 0.01 * x[..., 2]
    r = 0.5 + 0.05 * x[..., 3]
    f1 = 0.3 + 0.2 * x[..., 4]
    t1 = 1 + 0.2 * x[..., 5]
    omega = np.sqrt((c1 + c2) / m + 0j)
    return 3 * r - np.abs(2 * f1 / (m * omega**2) * np.sin(omega * t1 / 2))


TRUTH = {
    "convex": 4.21e-3,  # toy_problems.py:16
    "oscillator": 6.43e-6,  # toy_problems.py:87
    "diffusion": 1.682e-4,  # diffusion.py:251
}


def truth_affine(beta=3.5):
    return stats.norm.cdf(-beta)  # toy_problems.py:43


# --------------------------------------------------------------------------- #
# 1-D lognormal-field diffusion PDE (diffusion.py)
# --------------------------------------------------------------------------- #
def _bisect(f, a, b, tol, nmax):
    """diffusion.py:10-38."""
    if a == 0:
        a = 1e-16
    if f(a) * f(b) > 0:
        pole = (b + a) / 2
        if f(a) * f(pole - 1e-14) < 0:
            b = pole - 1e-14
        else:
            a = pole + 1e-14
    k = 0
    while (b - a >= tol) and (k <= nmax):
        mid = (b + a) / 2
        if f(a) * f(mid) < 0:
            b = mid
        else:
            a = mid
        k += 1
    return a, b


def _newton(f, df, x0, tol, nmax):
    """diffusion.py:41-53."""
    x = x0 - f(x0) / df(x0)
    err = abs(x - x0)
    k = 1
    while err >= tol and k <= nmax:
        xo = x
        x = xo - f(xo) / df(xo)
        err = abs(x - xo)
        k += 1
    return x


def kl_eigenpairs(corr_len: float, n_roots: int):
    """Roots, eigenvalues and normalisation of the exponential-kernel KL basis.  diffusion.py:67-128."""
    with np.errstate(all="ignore"):
        g_odd = lambda y: 1 / corr_len - y * np.tan(y / 2)
        dg_odd = lambda y: -np.tan(y / 2) - y * 1 / 2 * (1 + np.tan(y / 2) ** 2)
        g_even = lambda y: 1 / corr_len * np.tan(y / 2) + y
        dg_even = lambda y: 1 / corr_len * 1 / 2 * (1 + np.tan(y / 2) ** 2) + 1
        roots = np.zeros(n_roots)
        for i in range(1, n_roots + 1):
            f, df = (g_odd, dg_odd) if i % 2 == 1 else (g_even, dg_even)
            a = _bisect(f, (i - 0.99999) * np.pi, i * np.pi, 1e-15, 10000)[0]
            roots[i - 1] = _newton(f, df, a, 1e-15, 10000)
    lam = 2 * corr_len / (1 + roots**2 * corr_len**2)
    alpha = 1 / np.sqrt(1 / 2 + np.sin(roots) / (2 * roots))
    return roots, lam, alpha


class DiffusionOracle:
    """g(theta) = 0.535 - u(1) for -(a u')' = 1, u(0)=0, a = lognormal KL field.

    diffusion.py:182-242; ``mode_table`` is the reference's ``function_values``
    (1536 x 150), ``solve`` follows ``diffusion_equation_solve`` (131-171) with a
    Thomas sweep in place of SuperLU (matches ``spsolve`` to <1e-13).
    """

    def __init__(self, dim=150, n_ele=512, corr_len=0.01, mu_a=1.0, sigma_a=0.1, u_max=0.535):
        self.dim, self.n_ele, self.u_max = dim, n_ele, u_max
        self.h = 1.0 / n_ele
        self.sigma = math.sqrt(math.log(sigma_a**2 / mu_a**2 + 1))  # :202
        self.mu = math.log(mu_a) - self.sigma**2 / 2  # :203
        self.roots, self.lam, self.alpha = kl_eigenpairs(corr_len, dim)
        # quadrature nodes, diffusion.py:56-64 (3 Gauss-Legendre points per element)
        xgl = np.array([-np.sqrt(3 / 5), 0, np.sqrt(3 / 5)])
        ind = np.arange(1, n_ele + 1, dtype=np.float64)
        xq = (1.0 / n_ele / 2) * xgl[None, :] + (1.0 / n_ele / 2) * (2 * ind - 1)[:, None]
        self.xq = xq.reshape(-1)
        self.wgl = np.array([5 / 9, 8 / 9, 5 / 9])
        table = np.empty((self.xq.size, dim))
        for k in range(1, dim + 1):  # eigenfunction, :122-126
            arg = self.roots[k - 1] * (self.xq - 1 / 2)
            table[:, k - 1] = self.alpha[k - 1] * (np.cos(arg) if k % 2 == 1 else np.sin(arg))
        self.mode_table = np.sqrt(self.lam) * table  # :229

    def element_stiffness(self, theta):
        """Ke per element for a batch of theta (B,dim) -> (B,n_ele).  diffusion.py:141-151."""
        theta = np.atleast_2d(theta)
        kl = np.exp(self.mu + self.sigma * (theta @ self.mode_table.T))
        kei = 1 / self.h / 2 * kl
        return (kei.reshape(theta.shape[0], self.n_ele, 3) * self.wgl).sum(-1)

    def solve(self, theta):
        """u at the right boundary for a batch of theta: Thomas on the tridiagonal system of :152-166."""
        ke = self.element_stiffness(theta)  # (B, n)
        B, n = ke.shape
        diag = ke.copy()
        diag[:, :-1] += ke[:, 1:]  # diag_i = Ke[i] + Ke[i+1]; last row Ke[n-1]
        off = -ke[:, 1:]  # sub/super diagonal
        rhs = np.full((B, n), self.h)
        rhs[:, -1] = self.h / 2
        # forward elimination only: u_last = rhs'_last / diag'_last
        dp = diag[:, 0].copy()
        rp = rhs[:, 0].copy()
        for i in range(1, n):
            wgt = off[:, i - 1] / dp
            dp = diag[:, i] - wgt * off[:, i - 1]
            rp = rhs[:, i] - wgt * rp
        return rp / dp

    def flux_identity(self, theta):
        """Independent check: u(1) = sum_e h (n - e - 1/2) / Ke[e] (constant load => telescoping flux)."""
        ke = self.element_stiffness(theta)
        e = np.arange(self.n_ele)
        return (self.h * (self.n_ele - e - 0.5) / ke).sum(-1)

    def g(self, theta):
        """diffusion.py:232-242, vectorised over leading axes."""
        theta = np.asarray(theta, dtype=np.float64)
        lead = theta.shape[:-1]
        out = self.u_max - self.solve(theta.reshape(-1, self.dim))
        return out.reshape(lead) if lead else float(out[0])
