"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the CBREE hot path.

A NumPy/SciPy restatement of the reference's CBREE algorithm
(``/root/reference/src/rareeventestimation/solver.py`` and ``utilities.py``;
every function cites the lines it follows).  It exists so that the CUDA path
can be checked on the GPU box, where ``/root/reference`` is absent.

Pinning: ``tests/golden/make_golden.py`` runs the *live* reference in the build
container and commits per-iteration traces under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those traces (and, when
the reference is importable, against the reference directly).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (``cpu_baseline``
leg / ``--impl reference``) may import this module.  The product package
``rareeventestimation_b200`` never does.

Deliberate deviations from the reference (all <= 2e-14 relative):
  * ``softmax_weights`` sums with ``numpy.sum`` (pairwise); the reference's
    ``my_softmax`` uses Python's sequential builtin ``sum`` (utilities.py:127).
  * noise enters as an explicit ``Z`` (N,d) standard-normal block; the reference
    draws it inside ``Generator.multivariate_normal(method='svd')``
    (solver.py:667) -- ``x = mean + Z @ (U*sqrt(S)).T`` replays that bit-exactly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np
from scipy import optimize, stats

LOG_2PI = math.log(2.0 * math.pi)

TGT_KINDS = ("algebraic", "tanh", "arctan", "erf", "relu", "sigmoid")


# --------------------------------------------------------------------------- #
# N-vector primitives
# --------------------------------------------------------------------------- #
def log_target(g, e, sigma, kind="algebraic"):
    """log of the smoothed-indicator target density.  solver.py:321-363."""
    if kind == "tanh":  # :343-345
        return -np.log(2) + np.log1p(np.tanh(-sigma * g)) - e
    if kind == "arctan":  # :346-348
        return -np.log(2) + np.log1p(2 / np.pi * np.arctan(-sigma * g)) - e
    if kind == "algebraic":  # :349-355 (same operation order)
        return -np.log(2) + np.log1p(-sigma * g / np.sqrt(1 + sigma**2 * g**2)) - e
    if kind == "erf":  # :356-358
        return stats.norm.logcdf(-sigma * g) - e
    if kind == "relu":  # :359-360
        return -sigma * np.maximum(0, g) ** 3 - e
    return -np.log1p(np.exp(sigma * g)) - e  # sigmoid, :361-363


def softmax_weights(logw):
    """Normalised weights exp(logw - max) / sum.  utilities.py:105-142."""
    with np.errstate(under="ignore"):
        logw = np.asarray(logw, dtype=np.float64).squeeze()
        ok = np.isfinite(logw)
        lw = logw[ok]
        shifted = np.exp(lw - lw[np.argmax(lw)])
        out = np.zeros(ok.shape, dtype=np.float64)
        if np.any(shifted > 0.0):
            out[ok] = shifted
            return out / np.sum(out)
        # degenerate branches, utilities.py:128-140
        lo = np.amin(lw)
        if lo < 0:
            lw = lw + lo
        tot = np.sum(lw)
        if tot == 0:
            return out + 1 / len(out)
        out[ok] = lw
        return out / tot


def ess_of(weights):
    """(sum w)^2 / sum w^2.  solver.py:523, 596."""
    return np.sum(weights) ** 2 / np.sum(weights**2)


def log_cvar(log_samples, multiplier=None):
    """Coefficient of variation of exp(log_samples)*multiplier.  utilities.py:76-102."""
    ok = np.isfinite(log_samples)
    ls = log_samples[ok]
    if multiplier is None:
        multiplier = np.ones_like(ls)
    top = np.amax(ls)
    vals = np.exp(ls - top) * multiplier
    mean = np.average(vals)
    m = np.exp(top) * mean
    s = np.exp(top) * np.sqrt(np.average((vals - mean) ** 2))
    if m == 0:
        return np.inf
    out = s / m
    return np.inf if np.isnan(out) else out


def is_estimate(log_tgt, log_aux, g):
    """mean(exp(log_tgt - log_aux) * 1[g<=0]).  utilities.py:43-45."""
    return np.average(np.exp(log_tgt - log_aux) * (g <= 0))


def slope_of(y):
    """Least-squares slope of (i, y_i); NaN if any y is non-finite.  utilities.py:145-163."""
    y = np.atleast_1d(np.asarray(y, dtype=np.float64))
    if not np.all(np.isfinite(y)):
        return np.nan
    try:
        with np.errstate(all="ignore"):
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return np.polyfit(np.arange(len(y)), y, deg=1, full=True)[0][0]
    except Exception:
        return np.nan


# --------------------------------------------------------------------------- #
# (N,d) primitives
# --------------------------------------------------------------------------- #
def energy(X):
    """e_fun of a NormalProblem: -logpdf of N(0,I).  problem.py:112-113 -> utilities.py:70-73."""
    d = X.shape[-1]
    return -stats.multivariate_normal.logpdf(X, mean=np.zeros(d), cov=np.eye(d))


def energy_fast(X):
    """Closed form of :func:`energy` (0.5*|x|^2 + d/2*log 2pi); used where SciPy's eigh is too slow."""
    return 0.5 * np.einsum("...j,...j->...", X, X) + 0.5 * X.shape[-1] * LOG_2PI


def weighted_moments(X, w):
    """Weighted mean and ddof=0 covariance.  solver.py:633-636."""
    m = np.average(X, axis=0, weights=w)
    C = np.cov(X, aweights=w, ddof=0, rowvar=False)
    return m, np.atleast_2d(C)


def sample_moments(X):
    """Unweighted mean and ddof=1 covariance.  solver.py:670-671."""
    return np.average(X, axis=0), np.atleast_2d(np.cov(X, ddof=1, rowvar=False))


def noise_factor(c_noise):
    """F with  multivariate_normal(m, C, N) == m + Z @ F.T  (NumPy 'svd' method).  solver.py:667."""
    u, s, _ = np.linalg.svd(c_noise)
    return u * np.sqrt(s)


def gaussian_fit_logpdf(X, m, c):
    """log N(x; m, c) as SciPy evaluates it.  solver.py:672-674, 838-842."""
    return stats.multivariate_normal.logpdf(X, mean=m, cov=c)


# --------------------------------------------------------------------------- #
# per-iteration state
# --------------------------------------------------------------------------- #
@dataclass
class IterState:
    """Mirror of the reference's CBREECache fields (solver.py:63-116)."""

    X: np.ndarray
    g: np.ndarray
    e: np.ndarray
    m_w: Optional[np.ndarray] = None
    C_w: Optional[np.ndarray] = None
    sigma: float = 1.0
    beta: float = 1.0
    t: float = 0.5
    cvar_is: float = math.nan
    converged: bool = False
    iteration: int = 0
    slope_cvar: float = math.nan
    slope_sfp: float = math.nan
    estimate: float = 0.0
    est_uniform: float = 0.0
    est_sqrt: float = 0.0
    est_cvar: float = 0.0
    sfp: float = 0.0
    ess: float = math.nan
    n_evals: int = 0
    msg: str = ""
    # extras kept for the parity tests (not in the reference's cache)
    weights: Optional[np.ndarray] = None
    m_new: Optional[np.ndarray] = None
    c_new: Optional[np.ndarray] = None
    logq: Optional[np.ndarray] = None
    Z: Optional[np.ndarray] = None


@dataclass
class Options:
    """CBREE constructor options with the reference's defaults.  solver.py:176-200, 262-310."""

    stepsize_tolerance: float = 0.5
    t_step: float = -1
    num_steps: int = 100
    tgt_fun: str = "algebraic"
    observation_window: int = 5
    sigma_adaptivity: object = "cvar"
    beta_adaptivity: object = True
    cvar_tgt: float = 2
    sfp_tgt: float = 1 / 3
    ess_tgt: float = 1 / 2
    lip_sigma: float = 1
    divergence_check: bool = True
    convergence_check: bool = True
    save_history: bool = False
    fast_energy: bool = False  # oracle-only: closed-form e_fun for large N

    def initial_sigma(self):
        return 0 if isinstance(self.sigma_adaptivity, str) else self.sigma_adaptivity

    def sigma_mode(self):
        return self.sigma_adaptivity if isinstance(self.sigma_adaptivity, str) else "constant"

    def initial_beta(self):
        return 1 if self.beta_adaptivity is True else self.beta_adaptivity

    def beta_adaptive(self):
        return self.beta_adaptivity is True


class CbreeOracle:
    """CPU oracle of CBREE.solve().  solver.py:365-503 (GM / no-resample branch)."""

    def __init__(self, lsf: Callable, opts: Optional[Options] = None, rng=None, noise: Optional[Callable] = None):
        self.o = opts or Options()
        self._lsf = lsf
        self.n_evals = 0
        self.rng = rng if rng is not None else np.random.default_rng()
        # noise(N, d) -> Z block; default consumes the solver rng like solver.py:667 does
        self.noise = noise or (lambda n, d: self.rng.standard_normal((n, d)))

    # -- evaluation wrappers ------------------------------------------------- #
    def lsf(self, X):
        self.n_evals += int(np.prod(X.shape[:-1]))  # solver.py:376-378
        return self._lsf(X)

    def e_fun(self, X):
        return energy_fast(X) if self.o.fast_energy else energy(X)

    def logtgt(self, g, e, sigma):
        return log_target(g, e, sigma, self.o.tgt_fun)

    # -- weights ------------------------------------------------------------- #
    def compute_weights(self, st: IterState):
        """solver.py:614-640."""
        w = softmax_weights(self.logtgt(st.g, st.e, st.sigma) * st.beta)
        st.m_w, st.C_w = weighted_moments(st.X, w)
        st.weights = w
        if not np.isfinite(st.C_w).all():
            raise ValueError("Weighted covariance contains non-finite elements.")

    # -- sigma / beta -------------------------------------------------------- #
    def sigma_objective(self, st: IterState, sigma_new):
        """(cvar(exp(l(sigma_new) - l(sigma_old))) - cvar_tgt)^2.  solver.py:538-545."""
        delta = self.logtgt(st.g, 0.0, sigma_new) - self.logtgt(st.g, 0.0, st.sigma)
        return (log_cvar(delta) - self.o.cvar_tgt) ** 2

    def update_sigma_sfp(self, st: IterState):
        """solver.py:564-580."""
        cur = np.sum(st.g <= 0) / len(st.g)
        dmax = self.o.lip_sigma * np.log(1 / st.t)
        if self.o.sfp_tgt > np.finfo(float).eps:
            delta = np.sqrt(np.maximum(0, self.o.sfp_tgt - cur)) * dmax / np.sqrt(self.o.sfp_tgt)
        else:
            delta = 0
        st.sigma += delta

    def update_sigma_cvar(self, st: IterState):
        """solver.py:525-562."""
        try:
            if np.sum(st.g <= 0) / len(st.g) < self.o.sfp_tgt:
                self.update_sigma_sfp(st)
            else:
                res = optimize.minimize_scalar(
                    lambda s: self.sigma_objective(st, s),
                    bounds=[st.sigma, st.sigma + self.o.lip_sigma * np.log(1 / st.t)],
                    method="Bounded",
                )
                st.sigma = res.x
        except Exception:  # swallowed (logged only) like solver.py:557-562
            pass

    def ess_objective(self, st: IterState, beta):
        """ESS(beta) - ess_tgt*N.  solver.py:591-597."""
        w = softmax_weights(self.logtgt(st.g, st.e, st.sigma) * beta)
        return ess_of(w) - self.o.ess_tgt * len(w)

    def update_beta(self, st: IterState):
        """solver.py:582-612."""
        try:
            sol = optimize.root(lambda b: self.ess_objective(st, b), st.beta)
            if sol.x.item() > 0:
                st.beta = sol.x.item()
        except Exception:
            pass

    def update_beta_and_sigma(self, st: IterState):
        """solver.py:505-523."""
        mode = self.o.sigma_mode()
        if mode == "cvar":
            self.update_sigma_cvar(st)
        elif mode == "sfp":
            self.update_sigma_sfp(st)
        if self.o.beta_adaptive():
            self.update_beta(st)
        st.ess = ess_of(softmax_weights(self.logtgt(st.g, st.e, st.sigma) * st.beta))

    # -- the step ------------------------------------------------------------ #
    def perform_step(self, st: IterState, Z=None) -> IterState:
        """Euler-Maruyama step + IS density + weights of the new ensemble.  solver.py:664-694."""
        n, d = st.X.shape
        if Z is None:
            Z = self.noise(n, d)
        m_noise = (1 - st.t) * st.m_w
        c_noise = (1 - st.t**2) * (1 + st.beta) * st.C_w
        F = noise_factor(c_noise)
        X_new = st.t * st.X + (m_noise + Z @ F.T)
        m_new, c_new = sample_moments(X_new)
        logq = gaussian_fit_logpdf(X_new, m_new, c_new)
        new = IterState(X_new, self.lsf(X_new), self.e_fun(X_new))
        new.cvar_is = log_cvar(-new.e - logq, multiplier=(new.g <= 0))
        new.iteration = st.iteration + 1
        new.converged = st.converged
        new.sigma, new.beta, new.t = st.sigma, st.beta, st.t
        new.n_evals = self.n_evals
        new.m_new, new.c_new, new.logq, new.Z = m_new, c_new, logq, Z
        self.compute_weights(new)
        return new

    # -- step size ----------------------------------------------------------- #
    @staticmethod
    def _stack(m, c):
        return np.concatenate((m[None, ...], np.tril(c)))

    def initial_stepsize(self, st: IterState):
        """solver.py:783-817 (including the (f0 - f2/w) parenthesisation of :812)."""
        tol = self.o.stepsize_tolerance
        m0, c0 = sample_moments(st.X)
        x0 = self._stack(m0, c0)
        w = tol + tol * np.abs(x0)
        d0 = np.sqrt(np.average((x0 / w) ** 2))
        f0 = self._stack(-m0 + st.m_w, -2 * c0 + 2 * st.C_w)
        d1 = np.sqrt(np.average((f0 / w) ** 2))
        h0 = 0.01 * d0 / d1
        st.t = np.exp(-h0)
        self.update_beta_and_sigma(st)
        trial = self.perform_step(st)
        m2, c2 = sample_moments(trial.X)
        f2 = self._stack(-m2 + trial.m_w, -2 * c2 + 2 * trial.C_w)
        d2 = np.sqrt(np.average((f0 - f2 / w) ** 2)) / h0
        h1 = np.sqrt(0.01 / np.maximum(d1, d2))
        st.t = np.exp(-np.maximum(100 * h0, h1))
        st.n_evals = trial.n_evals

    def update_stepsize(self, states: List[IterState]):
        """Embedded two-stage error estimator.  solver.py:740-781."""
        s1, s2 = states[-2], states[-1]
        h = -np.log(s1.t)
        phi = (np.exp(-2 * h) - 1) / (-2 * h)
        bm1 = phi - 2 * (phi - 1) / (-2 * h)
        bm2 = 2 * (phi - 1) / (-2 * h)
        m1, c1 = sample_moments(s1.X)
        m2, c2 = sample_moments(s2.X)
        m2_hat = 2 * h * (bm1 * m1 + bm2 * m2)
        phi = (np.exp(-4 * h) - 1) / (-2 * h)
        bc1 = phi - 2 * (phi - 1) / (-2 * h)
        bc2 = 2 * (phi - 1) / (-2 * h)
        c2_hat = 2 * h * (bc1 * c1 + bc2 * c2)
        x, x_hat = self._stack(m2, c2), self._stack(m2_hat, c2_hat)
        tol = self.o.stepsize_tolerance
        w = tol + tol * np.maximum(np.abs(x), np.abs(x_hat))
        err = np.sqrt(np.average(((x - x_hat) / w) ** 2))
        q = 0.9 * np.sqrt(1 / err)
        s2.t = np.maximum(np.exp(-100), np.minimum(np.exp(-1e-5), np.exp(-q * h)))

    # -- stopping ------------------------------------------------------------ #
    def convergence_check(self, states: List[IterState]):
        """solver.py:696-738."""
        o = self.o
        last = states[-1]
        win = states[-o.observation_window:]
        sfp_counts = np.array([np.sum(s.g <= 0) for s in win])
        last.sfp = np.sum(last.g <= 0) / len(last.g)
        last.slope_cvar = slope_of(np.array([s.cvar_is for s in win]))
        last.slope_sfp = slope_of(sfp_counts)
        mode = o.sigma_mode()
        if mode == "cvar":
            last.converged = bool((last.cvar_is <= o.cvar_tgt) and o.convergence_check)
            if last.converged:
                last.msg += "Converged with given `cvar_tgt`."
            if o.divergence_check and last.iteration >= o.observation_window:
                sfp_mean = np.average([s.sfp for s in win])
                last.converged = bool(last.converged or (last.slope_cvar > 0.0 and sfp_mean >= o.sfp_tgt))
                if last.converged:
                    last.msg += "Converged due to `divergence_check`."
        if mode == "sfp":
            last.converged = bool((last.sfp >= o.sfp_tgt) and o.convergence_check)
            if last.converged:
                last.msg += "Converged with given `sfp_tgt`."

    # -- final estimates ----------------------------------------------------- #
    def importance_sampling(self, st: IterState):
        """solver.py:819-846 (GM branch): refit N(m', c') to the ensemble, IS estimate."""
        m, c = sample_moments(st.X)
        aux = gaussian_fit_logpdf(st.X, m, c)
        tgt = -self.e_fun(st.X)  # gaussian_logpdf, utilities.py:70-73
        st.estimate = is_estimate(tgt, aux, st.g)

    def weighted_estimates(self, states: List[IterState]):
        """solver.py:848-864."""
        k = min(len(states), self.o.observation_window)
        sqrt_w = np.sqrt(np.arange(k))
        with np.errstate(all="ignore"):
            cvar_w = 1 / np.array([s.cvar_is**2 for s in states[-k:]])
        cvar_w[~np.isfinite(cvar_w)] = 0.0
        pfs = [s.estimate for s in states[-k:]]
        last = states[-1]
        last.est_uniform = np.average(pfs)
        last.est_cvar = np.average(pfs, weights=cvar_w) if np.sum(cvar_w) > 0 else 0.0
        last.est_sqrt = np.average(pfs, weights=sqrt_w)  # raises for k == 1, like the reference

    # -- main loop ----------------------------------------------------------- #
    def solve(self, X0, keep_all=False):
        """solver.py:365-503.  Returns (states, msg); states are the retained caches."""
        o = self.o
        X0 = np.asarray(X0, dtype=np.float64)
        self.n_evals = 0
        st = IterState(X0, self.lsf(X0), self.e_fun(X0), sigma=o.initial_sigma(), beta=o.initial_beta())
        st.n_evals = self.n_evals
        self.compute_weights(st)
        if o.t_step > 0:
            st.t = o.t_step
        else:
            self.initial_stepsize(st)
        states = [st]
        history = [st]
        msg = "Success"
        while not states[-1].converged and states[-1].iteration <= o.num_steps:
            if o.t_step <= 0 and len(states) > 1 and states[-1].iteration % 2 == 0:
                self.update_stepsize(states)
            self.update_beta_and_sigma(states[-1])
            try:
                new = self.perform_step(states[-1])
            except Exception as exc:
                msg = str(exc)
                break
            states.append(new)
            history.append(new)
            if not o.save_history and len(states) > o.observation_window:
                states.pop(0)
            self.convergence_check(states)
        if not states[-1].converged and states[-1].iteration > o.num_steps:
            msg = "Not Converged."
        for s in states:
            self.importance_sampling(s)
        self.weighted_estimates(states)
        return (history if keep_all else states), msg
