"""Host-array helpers with the reference's names (``utilities.py``): user callbacks call them
(``ree.my_log_cvar`` in docs/benchmarking/benchmarking-scripts/dimension_study.py:20-22).  The N-sized reductions
run in the CUDA kernels (``ree_reduce_scaled``, ``ree_energy``); there is no CPU path behind them."""
from __future__ import annotations

import math

import numpy as np

from .engine import Engine
from .problem import StandardNormalEnergy
from .solver import get_slope  # noqa: F401  (utilities.py:145-163)


def _exp_sums(ell: np.ndarray):
    """[max, sum exp(l - max), sum exp(2 (l - max)), #finite] of a host vector on the device."""
    eng = Engine.get()
    ell = np.ascontiguousarray(ell, dtype=np.float64).reshape(-1)
    n = ell.shape[0]
    if n == 0:
        return np.array([-math.inf, 0.0, 0.0, 0.0])
    ld = max(8, (n + 7) // 8 * 8)
    return eng.fetch(eng.reduce_scaled(eng.upload_vector(ell, ld), n, 1, 1.0))


def gaussian_logpdf(x: np.ndarray) -> np.ndarray:
    """log N(x; 0, I) along the last axis (utilities.py:70-73)."""
    return -np.asarray(StandardNormalEnergy()(x))


def my_log_cvar(log_samples: np.ndarray, multiplier=None) -> float:
    """Coefficient of variation of ``exp(log_samples) * multiplier`` over the finite entries of ``log_samples``
    (utilities.py:76-102): inf if the mean is zero or the result is NaN.  ``multiplier`` must be non-negative."""
    ls = np.asarray(log_samples, dtype=np.float64).reshape(-1)
    fin = np.isfinite(ls)
    cnt = float(np.count_nonzero(fin))
    if cnt == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    if multiplier is not None:
        m = np.asarray(multiplier, dtype=np.float64).reshape(-1)
        if np.any(m[fin] < 0):
            raise ValueError("my_log_cvar: the multiplier must be non-negative")
        with np.errstate(divide="ignore"):
            ls = np.where(fin & (m > 0), ls + np.log(np.where(m > 0, m, 1.0)), -math.inf)
    t = _exp_sums(ls)
    mean = t[1] / cnt
    if mean == 0:
        return math.inf
    out = math.sqrt(max(t[2] / cnt - mean * mean, 0.0)) / mean
    return math.inf if math.isnan(out) else out


def my_softmax(weights_in: np.ndarray) -> np.ndarray:
    """exp(w - max w) / sum over the finite entries, zero elsewhere (utilities.py:105-142); the normalising sums come
    from the device reduction."""
    w = np.asarray(weights_in, dtype=np.float64).reshape(-1)
    t = _exp_sums(w)
    if t[3] <= 0:
        raise ValueError("attempt to get argmax of an empty sequence")
    with np.errstate(all="ignore"):
        out = np.where(np.isfinite(w), np.exp(w - t[0]), 0.0)
    s = t[1]
    if not (s > 0 and math.isfinite(s)):  # degenerate sums: all mass on the largest entry (utilities.py:128-140)
        out = np.zeros_like(w)
        out[int(np.nanargmax(np.where(np.isfinite(w), w, -math.inf)))] = 1.0
        return out
    return out / s


def importance_sampling(target_pdf_evals, aux_pdf_evals, lsf_evals, logpdf=False) -> float:
    """mean(target / aux * 1[lsf <= 0]) (utilities.py:29-48); with ``logpdf`` the inputs are log-densities."""
    tgt = np.asarray(target_pdf_evals, dtype=np.float64).reshape(-1)
    aux = np.asarray(aux_pdf_evals, dtype=np.float64).reshape(-1)
    g = np.asarray(lsf_evals, dtype=np.float64).reshape(-1)
    with np.errstate(all="ignore"):
        lw = (tgt - aux) if logpdf else (np.log(tgt) - np.log(aux))
    t = _exp_sums(np.where(g <= 0, lw, -math.inf))
    if not (t[1] > 0) or not math.isfinite(t[0]):
        return 0.0
    return math.exp(t[0]) * t[1] / g.shape[0]
