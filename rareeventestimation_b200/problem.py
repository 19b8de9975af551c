"""Problem formulation -- same protocol as the reference's ``problem/problem.py``:
``prob.lsf``, ``prob.e_fun``, ``prob.sample`` (N,d), ``prob.set_sample``, ``prob.prob_fail_true``,
``prob.name``, ``prob.eranataf_dist``.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
from scipy import special

from .engine import Engine

_STREAM_SAMPLE = 0x53 << 24  # Philox draw index reserved for initial samples


class StandardNormalEnergy:
    """e_fun of a NormalProblem: -log N(x; 0, I) = |x|^2/2 + d/2 log 2pi (problem.py:112-113 ->
    utilities.py:70-73).  Callable on host arrays; evaluated on the GPU (``ree_energy``)."""

    is_standard_normal = True

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        d = x.shape[-1]
        lead = x.shape[:-1]
        flat = np.ascontiguousarray(x.reshape(-1, d))
        eng = Engine.get()
        ens = eng.upload(flat)
        out = eng.download_vector(eng.energy(ens), ens.n)
        return out.reshape(lead) if lead else float(out[0])


class StandardNormalNataf:
    """The slice of ``era/ERANataf.py`` a NormalProblem needs: independent standard-normal marginals.
    ``random`` follows ERANataf.random (372-386): a dimension-major N(0,1) block from
    ``default_rng(seed)``, then the per-marginal ``ppf(cdf(.))`` round trip of the Nataf map."""

    def __init__(self, dim: int):
        self.dim = int(dim)
        self.Marginals = [None] * self.dim
        self.Rho_X = np.eye(self.dim)
        self.A = np.eye(self.dim)

    def random(self, n=1, seed=None):
        n = int(n)
        U = np.random.default_rng(seed).standard_normal((self.dim, n))
        out = np.empty((n, self.dim))
        for j in range(self.dim):
            out[:, j] = special.ndtri(special.ndtr(U[j]))
        return np.squeeze(out)

    def U2X(self, U, Jacobian=False):
        U = np.array(U, ndmin=2, dtype=np.float64)
        X = special.ndtri(special.ndtr(U))
        return (np.squeeze(X), np.eye(self.dim)) if Jacobian else np.squeeze(X)

    def X2U(self, X, Jacobian=False):
        return self.U2X(X, Jacobian)

    def pdf(self, X):
        X = np.array(X, ndmin=2, dtype=np.float64)
        return np.exp(-0.5 * np.sum(X * X, axis=-1) - 0.5 * self.dim * np.log(2 * np.pi))


class DeviceSample:
    """An initial sample that was drawn on the device (Philox) and never visited the host.
    Quacks like the (N,d) ndarray the reference stores in ``prob.sample``."""

    def __init__(self, n: int, d: int, seed: int, draw: int = _STREAM_SAMPLE):
        self.shape = (int(n), int(d))
        self.seed = int(seed)
        self.draw = int(draw)  # Philox draw index of the stream this sample is a window of
        self.ndim = 2
        self.dtype = np.dtype(np.float64)

    def materialise(self, engine: Engine, lo: int = 0, hi: Optional[int] = None):
        """Ensemble for the global particle range [lo, hi) (rank-invariant counters)."""
        hi = self.shape[0] if hi is None else hi
        return engine.philox_normal(hi - lo, self.shape[1], self.seed, self.draw, first=lo)

    def __array__(self, dtype=None, copy=None):
        eng = Engine.get()
        return eng.download(self.materialise(eng))

    def __len__(self):
        return self.shape[0]


class ShardedHostSample:
    """This rank's rows ``[lo, lo + len(local))`` of a global (n_total, d) host sample (one process per GPU:
    no rank has to hold the other ranks' particles in host memory)."""

    def __init__(self, local: np.ndarray, n_total: int, lo: int):
        self.local = local
        self.lo = int(lo)
        self.shape = (int(n_total), int(local.shape[1]))
        self.ndim = 2
        self.dtype = np.dtype(np.float64)

    def __len__(self):
        return self.shape[0]


class Problem:
    """A rare event estimation problem (problem.py:11-74)."""

    def __init__(self, lsf: Callable, e_fun: Callable, sample, sample_gen=None, prob_fail_true=None, mpp=None,
                 eranataf_dist=None, hints=None, name=None):
        self.lsf = lsf
        self.e_fun = e_fun
        self.sample = sample
        self.prob_fail_true = prob_fail_true
        self.mpp = mpp
        if sample_gen is not None:
            self.sample_gen = sample_gen
        if eranataf_dist is not None:
            self.eranataf_dist = eranataf_dist
        self.hints = hints if isinstance(hints, dict) else {}
        self.name = f"Unnamed Problem at {id(self)}" if name is None else name

    def set_sample(self, arg, seed=None, device=False):
        """Reset ``self.sample`` (problem.py:57-74).  ``device=True`` draws the sample with Philox in HBM
        (no host round trip) instead of the reference's host generator."""
        if isinstance(arg, (int, np.integer)):
            assert hasattr(self, "sample_gen"), "Problem has no sample generator. Provide a sample as a ndarray!"
            if device:
                if not is_standard_normal_problem(self):
                    raise ValueError("set_sample(device=True) draws N(0, I) on the device: only for problems on the "
                                     "standard normal space (NormalProblem); this problem has its own sample_gen")
                d = self.sample.shape[-1]
                self.sample = DeviceSample(int(arg), d, 0 if seed is None else int(seed))
            else:
                self.sample = self.sample_gen(int(arg), seed=seed)
        elif isinstance(arg, (np.ndarray, DeviceSample, ShardedHostSample)) and arg.shape[-1] == self.sample.shape[-1]:
            self.sample = arg
        else:
            print("arg is no integer or sample of correct dimension.")
        return self


class NormalProblem(Problem):
    """LSF on the standard normal space (problem.py:77-130): energy and sampler are known."""

    def __init__(self, lsf: Callable, dim: int, sample_size: int, prob_fail_true=None, mpp=None, hints=None,
                 name=None):
        dist = StandardNormalNataf(dim)

        def sample_gen(sample_size, seed=None):
            return dist.random(n=sample_size, seed=seed).reshape(int(sample_size), dim)

        super().__init__(lsf, StandardNormalEnergy(), sample_gen(sample_size), sample_gen=sample_gen,
                         prob_fail_true=prob_fail_true, mpp=mpp, eranataf_dist=dist, hints=hints, name=name)


def is_standard_normal_problem(prob) -> bool:
    """True if ``prob`` lives on the standard normal space: its sampler draws N(0, I), its energy is
    -log N(x; 0, I) and SIS needs no U2X transform (solver.py:1144 ``transform_lsf=not isinstance(prob, NormalProblem)``)."""
    return isinstance(prob, NormalProblem) or bool(getattr(prob.e_fun, "is_standard_normal", False))


class Vectorizer:
    """Dispatching wrapper around a point-wise LSF (problem.py:133-198): a 1-d point, a 2-d array of
    rows, or a d-tuple of meshgrid arrays.  Counts evaluations in ``num_evals``."""

    num_evals = 0

    def __init__(self, lsf: Callable, d: int, lsf_2d=None, lsf_msh=None) -> None:
        if isinstance(lsf, Vectorizer):
            self.d, self.lsf_1d, self.lsf_2d, self.lsf_msh = lsf.d, lsf.lsf_1d, lsf.lsf_2d, lsf.lsf_msh
        else:
            self.d = d
            self.lsf_1d = lsf
            self.lsf_2d = lsf_2d if lsf_2d is not None else (lambda xx: np.apply_along_axis(lsf, 1, xx))
            if lsf_msh is None:

                def lsf_msh(*xi):
                    out = np.zeros(xi[0].shape)
                    for idx, _ in np.ndenumerate(out):
                        out[idx] = lsf(np.array([x[idx] for x in xi]))
                    return out

            self.lsf_msh = lsf_msh
        self.num_evals = 0

    def __call__(self, *args):
        if len(args) == 1 and args[0].shape == (self.d,):
            self.num_evals += 1
            return self.lsf_1d(args[0])
        if len(args) == 1 and args[0].ndim == 2 and args[0].shape[1] == self.d:
            self.num_evals += args[0].shape[0]
            return self.lsf_2d(args[0])
        if len(args) == self.d and all(x.ndim == self.d and x.shape == args[0].shape for x in args):
            self.num_evals += int(np.prod(args[0].shape))
            return self.lsf_msh(*args)
        raise ValueError("Vectorizer: arguments do not match the LSF's dimension")
