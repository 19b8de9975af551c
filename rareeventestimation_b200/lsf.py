"""Limit-state functions with a device descriptor.

Each class is callable on host arrays ``(..., d) -> (...)`` like the reference's LSFs
(``problem/toy_problems.py``, ``problem/diffusion.py``) -- the evaluation itself runs on the GPU
through ``ree_lsf_eval`` -- and hands the solvers a :class:`~.engine.DeviceLsf` so that the fused
kernels evaluate it without leaving HBM.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from . import _lib
from .engine import DeviceLsf, Engine


class DeviceLSF:
    """Base class of limit-state functions the engine can evaluate on the device."""

    kind = _lib.LSF_EXTERNAL
    dim: Optional[int] = None

    def descriptor(self, engine: Engine, d: int) -> DeviceLsf:
        raise NotImplementedError

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        d = x.shape[-1]
        lead = x.shape[:-1]
        flat = np.ascontiguousarray(x.reshape(-1, d))
        if flat.shape[0] == 0:
            return np.zeros(lead)
        eng = Engine.get()
        ens = eng.upload(flat)
        g = eng.lsf_eval(self.descriptor(eng, d), ens)
        out = eng.download_vector(g, ens.n)
        return out.reshape(lead) if lead else float(out[0])


class AffineLSF(DeviceLSF):
    """g(x) = offset + <coef, x>; ``coef=None`` is the reference's linear problem
    ``-sum(x)/sqrt(d) + beta`` (toy_problems.py:35-48, README.md:31-49)."""

    kind = _lib.LSF_AFFINE

    def __init__(self, offset=3.5, coef=None):
        self.offset = float(offset)
        self.coef = None if coef is None else np.asarray(coef, dtype=np.float64)
        self.dim = None if coef is None else int(self.coef.shape[0])
        self._dev = {}

    def descriptor(self, engine, d):
        if self.coef is None:
            return DeviceLsf(self.kind, d, None, self.offset)
        if d != self.dim:
            raise ValueError(f"AffineLSF has {self.dim} coefficients, got d={d}")
        key = engine.device.index
        if key not in self._dev:
            self._dev[key] = engine.to_device(self.coef)
        return DeviceLsf(self.kind, d, self._dev[key], self.offset)

    def __deepcopy__(self, memo):
        return AffineLSF(self.offset, None if self.coef is None else self.coef.copy())


class ConvexLSF(DeviceLSF):
    """toy_problems.py:12-13."""

    kind = _lib.LSF_CONVEX
    dim = 2

    def descriptor(self, engine, d):
        return DeviceLsf(self.kind, d)


class FujitaRackwitzLSF(DeviceLSF):
    """g(x) = -sum_j log Phi(-x_j) - c   (toy_problems.py:54-66)."""

    kind = _lib.LSF_FUJITA_RACKWITZ

    def __init__(self, c):
        self.c = float(c)

    def descriptor(self, engine, d):
        return DeviceLsf(self.kind, d, None, self.c)


class OscillatorLSF(DeviceLSF):
    """Nonlinear oscillator, d = 6 (toy_problems.py:72-80)."""

    kind = _lib.LSF_OSCILLATOR
    dim = 6

    def descriptor(self, engine, d):
        if d != 6:
            raise ValueError("the nonlinear oscillator needs d = 6")
        return DeviceLsf(self.kind, 6)


class DiffusionLSF(DeviceLSF):
    """1-D diffusion equation with a lognormal KL-expanded coefficient (problem/diffusion.py).

    The KL mode table (``function_values``, diffusion.py:222-229: sqrt(lambda_k) phi_k(x_q) at the
    3 Gauss-Legendre points of each of the ``n_ele`` elements) is set up once on the host -- it is
    150 scalar root finds, not N-sized work -- and lives in HBM afterwards."""

    kind = _lib.LSF_DIFFUSION

    def __init__(self, dim=150, n_ele=512, corr_len=0.01, mu_a=1.0, sigma_a=0.1, u_max=0.535):
        self.dim, self.n_ele, self.u_max = int(dim), int(n_ele), float(u_max)
        self.sigma = math.sqrt(math.log(sigma_a**2 / mu_a**2 + 1))  # diffusion.py:202
        self.mu = math.log(mu_a) - self.sigma**2 / 2  # diffusion.py:203
        self.corr_len = corr_len
        self._table = None
        self._dev = {}

    # -- KL basis of the exponential covariance kernel on [0,1] (diffusion.py:67-128) -- #
    @staticmethod
    def _kl_roots(corr_len, n_roots):
        inv = 1.0 / corr_len
        roots = np.empty(n_roots)
        with np.errstate(all="ignore"):
            for i in range(1, n_roots + 1):
                if i % 2 == 1:
                    f = lambda y: inv - y * np.tan(y / 2)
                    df = lambda y: -np.tan(y / 2) - y / 2 * (1 + np.tan(y / 2) ** 2)
                else:
                    f = lambda y: inv * np.tan(y / 2) + y
                    df = lambda y: inv / 2 * (1 + np.tan(y / 2) ** 2) + 1
                a, b = (i - 0.99999) * np.pi, i * np.pi
                if f(a) * f(b) > 0:  # bracket across the pole of tan (diffusion.py:17-23)
                    pole = (a + b) / 2
                    if f(a) * f(pole - 1e-14) < 0:
                        b = pole - 1e-14
                    else:
                        a = pole + 1e-14
                it = 0
                while b - a >= 1e-15 and it <= 10000:
                    mid = (a + b) / 2
                    if f(a) * f(mid) < 0:
                        b = mid
                    else:
                        a = mid
                    it += 1
                y = a - f(a) / df(a)
                step, it = abs(y - a), 1
                while step >= 1e-15 and it <= 10000:
                    y_old = y
                    y = y_old - f(y_old) / df(y_old)
                    step = abs(y - y_old)
                    it += 1
                roots[i - 1] = y
        return roots

    def mode_table(self) -> np.ndarray:
        if self._table is None:
            w = self._kl_roots(self.corr_len, self.dim)
            lam = 2 * self.corr_len / (1 + w**2 * self.corr_len**2)
            alpha = 1 / np.sqrt(0.5 + np.sin(w) / (2 * w))
            half = 0.5 / self.n_ele
            nodes = np.array([-math.sqrt(3 / 5), 0.0, math.sqrt(3 / 5)])
            centers = half * (2 * np.arange(1, self.n_ele + 1) - 1)
            xq = (half * nodes[None, :] + centers[:, None]).reshape(-1)
            arg = (xq[:, None] - 0.5) * w[None, :]
            odd = (np.arange(1, self.dim + 1) % 2 == 1)[None, :]
            self._table = np.sqrt(lam)[None, :] * alpha[None, :] * np.where(odd, np.cos(arg), np.sin(arg))
        return self._table

    def descriptor(self, engine, d):
        if d != self.dim:
            raise ValueError(f"diffusion LSF has {self.dim} modes, got d={d}")
        key = engine.device.index
        if key not in self._dev:
            self._dev[key] = engine.to_device(self.mode_table())
        return DeviceLsf(self.kind, d, self._dev[key], self.u_max, self.mu, self.sigma, self.n_ele)

    def __deepcopy__(self, memo):
        new = DiffusionLSF.__new__(DiffusionLSF)
        new.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_dev"})
        new._dev = {}
        return new
