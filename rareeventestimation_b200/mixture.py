"""von Mises-Fisher-Nakagami mixture with one component on the engine.

Mirror of the reference's ``VMFNMixture`` (mixturemodel.py:153-239) as CBREE uses it (``VMFNMixture(1)``,
solver.py:654-663, 828-836; ``callback_vmfnm`` in docs/benchmarking/benchmarking-scripts/dimension_study.py:12-25):
``fit(sample, weights_sample=None)``, ``sample(N, rng)``, ``logpdf(sample)``, ``pdf(sample)`` on host ``(N, d)``
arrays, with the fitted parameters in the reference's shapes (``mu (d,1)``, ``kappa (1,)``, ``m (1,1)``,
``omega (1,1)``, ``alpha (1,)``).  The N-sized work runs in ``ree_vmfn_fit_sums`` / ``ree_vmfn_sample`` /
``ree_vmfn_logpdf`` (csrc/ree_vmfn.cu); the ``*_device`` methods are what the solver calls on HBM-resident ensembles.
More than one component (the EM proper, era/EMvMFNM.py:44-93) is not built: CBREE never asks for it.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
from scipy.special import gammaln, iv

from .engine import Engine, Ensemble


class MixtureModel:
    """Parent of the mixture models (mixturemodel.py:27-40); ``fitted`` is what CBREE's importance sampling checks."""

    def __init__(self, num_comps: int) -> None:
        self.num_comps = num_comps
        self.fitted = False


def _house(x: np.ndarray):
    """Householder vector v and factor b with (I - b v v^T) e_d = x/|x| (EMvMFNM.py:401-429)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    n = x.size
    s = float(x[: n - 1] @ x[: n - 1])
    v = np.concatenate([x[: n - 1], [1.0]])
    if s == 0.0:
        return v, 0.0
    nrm = math.sqrt(x[n - 1] * x[n - 1] + s)
    v[n - 1] = x[n - 1] - nrm if x[n - 1] <= 0 else -s / (x[n - 1] + nrm)
    b = 2.0 * v[n - 1] * v[n - 1] / (s + v[n - 1] * v[n - 1])
    return v / v[n - 1], b


def _log_bessel_i(nu: float, x: float) -> float:
    """log I_nu(x); uniform asymptotic expansion (Abramowitz & Stegun 9.7.7) for nu != 0 (EMvMFNM.py:239-252)."""
    if nu == 0:
        return math.log(iv(nu, x))
    frac = x / nu
    square = 1.0 + frac * frac
    root = math.sqrt(square)
    eta = root + math.log(frac) - math.log(1.0 + root)
    return -math.log(math.sqrt(2.0 * math.pi * nu)) + nu * eta - 0.25 * math.log(square)


class VMFNMixture(MixtureModel):
    """Fit, sample and evaluate a one-component von Mises-Fisher-Nakagami model."""

    def __init__(self, num_comps: int) -> None:
        super().__init__(num_comps)
        if num_comps != 1:
            raise NotImplementedError("only VMFNMixture(1) is built (what CBREE uses, solver.py:659, 829)")

    # ------------------------------------------------------------------ #
    # parameters
    # ------------------------------------------------------------------ #
    def _set_from_sums(self, sums: np.ndarray, d: int) -> None:
        """M-step of EMvMFNM.py:169-199 for one component (responsibilities identically 1)."""
        S, s2, s4, sw = sums[:d], float(sums[d]), float(sums[d + 1]), float(sums[d + 2])
        norm_mu = float(np.sqrt(np.sum(S * S)))
        if norm_mu == 0:
            print("error")
        self.mu = (S / norm_mu).reshape(d, 1)
        xi = min(norm_mu / sw, 0.95)
        self.kappa = np.array([(xi * d - xi**3) / (1 - xi**2)])
        omega = np.float64(s2 / sw)
        mu4 = np.float64(s4 / sw)
        with np.errstate(all="ignore"):
            m = float(omega**2 / (mu4 - omega**2))  # +-inf / nan for a degenerate radius law -> d/2 below
        if not (m >= 0) or m > 20 * d:
            m = d / 2
        self.m = np.array([[m]])
        self.omega = np.array([[float(omega)]])
        self.alpha = np.array([1.0])
        self.fitted = True
        self.num_comps = 1
        self._d = d
        self._dev_mu = None

    def _scalars(self):
        return float(self.kappa[0]), float(self.m[0, 0]), float(self.omega[0, 0])

    def _c0(self) -> float:
        """vMF normalisation (EMvMFNM.py:208-224) + the constant part of the Nakagami density (229-236)."""
        d = self._d
        kappa, m, om = self._scalars()
        if kappa == 0:
            c_vmf = -(math.log(d) + math.log(math.pi ** (d / 2)) - gammaln(d / 2 + 1))
        elif kappa > 0:
            c_vmf = (d / 2 - 1) * math.log(kappa) - (d / 2) * math.log(2 * math.pi) - _log_bessel_i(d / 2 - 1, kappa)
        else:
            raise ValueError("Concentration parameter kappa must not be negative!")
        return c_vmf + math.log(2) + m * (math.log(m) - math.log(om)) - float(gammaln(m))

    def _mu_device(self, eng: Engine):
        if self._dev_mu is None:
            self._dev_mu = eng.to_device_small(np.ascontiguousarray(self.mu[:, 0]))
        return self._dev_mu

    # ------------------------------------------------------------------ #
    # device paths (what CBREE calls)
    # ------------------------------------------------------------------ #
    def fit_device(self, eng: Engine, ens: Ensemble, w=None) -> None:
        self._set_from_sums(eng.fetch_array(eng.vmfn_fit_sums(ens, w)), ens.d)

    def sample_device(self, eng: Engine, dst: Ensemble, seed: int, draw: int, variates=None) -> Ensemble:
        """Fill ``dst`` with samples.  ``variates`` = (r, w, V) host arrays of *global* size replays injected draws
        (parity mode, see :func:`vmfn_variates_numpy`); otherwise Philox streams keyed by the global particle index."""
        assert self.fitted, "Fit vMFNM before accessing parameters!"
        kappa, m, om = self._scalars()
        v, b = _house(self.mu[:, 0])
        hv = eng.to_device_small(v)
        if variates is None:
            return eng.vmfn_sample(dst, hv, b, kappa, m, om, seed, draw)
        r, w, V = variates
        lo, hi = dst.first, dst.first + dst.n
        rd, wd = eng.upload_vector(r[lo:hi], dst.ld), eng.upload_vector(w[lo:hi], dst.ld)
        if dst.d > 1 and V.shape[1] > 0:
            vd = eng.upload(np.ascontiguousarray(V[lo:hi]), first=dst.first)
        else:
            vd = eng.alloc_ensemble(dst.n, 1, dst.first)
        if vd.ld != dst.ld:
            raise ValueError("leading dimensions differ")
        return eng.vmfn_sample(dst, hv, b, kappa, m, om, 0, 0, r=rd, w=wd, v=vd)

    def logpdf_device(self, eng: Engine, ens: Ensemble, logq=None, is4=None, fuse_lsf=None) -> None:
        assert self.fitted, "Fit vMFNM before accessing parameters!"
        kappa, m, om = self._scalars()
        eng.vmfn_logpdf(ens, self._mu_device(eng), kappa, m, om, self._c0(), logq=logq, is4=is4, fuse_lsf=fuse_lsf)

    # ------------------------------------------------------------------ #
    # the reference's host-array interface
    # ------------------------------------------------------------------ #
    def fit(self, sample: np.ndarray, weights_sample: Optional[np.ndarray] = None) -> None:
        """Fit to a host sample (N, d); ``weights_sample`` (N,) defaults to uniform (mixturemodel.py:165-180)."""
        eng = Engine.get()
        ens = eng.upload(np.asarray(sample, dtype=np.float64))
        w = None
        if weights_sample is not None:
            w = eng.upload_vector(np.asarray(weights_sample, dtype=np.float64).reshape(-1), ens.ld)
        self.fit_device(eng, ens, w)

    def sample(self, N: int, rng=None) -> np.ndarray:
        """(N, d) array of samples (mixturemodel.py:182-194).  The variates are Philox streams seeded from ``rng``."""
        assert self.fitted, "Fit vMFNM before accessing parameters!"
        eng = Engine.get()
        rng = np.random.default_rng() if rng is None else rng
        seed = int(rng.integers(0, 2**63 - 1))
        dst = eng.alloc_ensemble(N, self._d, 0, with_vectors=False)
        self.sample_device(eng, dst, seed, 0)
        return eng.download(dst)

    def logpdf(self, sample: np.ndarray) -> np.ndarray:
        """log-density at the rows of ``sample`` (mixturemodel.py:196-215)."""
        assert self.fitted, "Fit vMFNM before accessing parameters!"
        eng = Engine.get()
        ens = eng.upload(np.asarray(sample, dtype=np.float64))
        out = eng.empty(ens.ld)
        self.logpdf_device(eng, ens, logq=out)
        return eng.download_vector(out, ens.n)

    def pdf(self, sample: np.ndarray) -> np.ndarray:
        return np.exp(self.logpdf(sample))

    def predict(self, sample: np.ndarray) -> np.ndarray:
        """Component every row belongs to (mixturemodel.py:226-239): with one component, component 0."""
        assert self.fitted, "Fit vMFNM before accessing parameters!"
        return np.zeros(np.asarray(sample).shape[0], dtype=np.int64)


def vmfn_variates_numpy(rng, n: int, d: int, kappa: float, m: float, omega: float):
    """Parity mode (``CBREE(noise="numpy")``): consume ``rng`` exactly like ``vMFNM_sample`` does for one component
    (EMvMFNM.py:274-285, 324-353) and return the *primitive* variates -- radii, accepted cosines, raw normals of the
    complement direction.  Only the scalar accept test runs here (it decides how many draws each sample consumes);
    normalisation, scaling and the Householder reflection of the N x d array happen in ``ree_vmfn_sample``."""
    r = np.sqrt(rng.gamma(m, scale=omega / m, size=(n, 1))).reshape(-1)
    t1 = np.sqrt(4 * kappa * kappa + (d - 1) * (d - 1))
    b = (-2 * kappa + t1) / (d - 1)
    x0 = (1 - b) / (1 + b)
    mh = (d - 1) / 2
    c = kappa * x0 + (d - 1) * np.log(1 - x0 * x0)
    w = np.empty(n)
    V = np.empty((n, d - 1))
    for i in range(n):
        t, u = -1000.0, 1.0
        while t < np.log(u):
            z = rng.beta(mh, mh)
            u = rng.uniform()
            wi = (1 - (1 + b) * z) / (1 - (1 - b) * z)
            t = kappa * wi + (d - 1) * np.log(1 - x0 * wi) - c
        w[i] = wi
        V[i] = rng.normal(size=(1, d - 1))
    return r, w, V
