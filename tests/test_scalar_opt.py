"""The scalar searches the persistent sigma / beta kernel runs (csrc/scalar_opt.cuh) against SciPy, on the CPU.

``ree_scalar_brent`` / ``ree_scalar_hybr1`` are host builds of the same source the kernel compiles; the reference
uses ``scipy.optimize.minimize_scalar(method="Bounded")`` (solver.py:548) and ``scipy.optimize.root`` (solver.py:601).
The sequence of trial points and the result must agree bit for bit."""
import ctypes as C

import numpy as np
import pytest
from scipy.optimize import minimize_scalar, root

from rareeventestimation_b200 import _lib


def _brent(fun, lo, hi, xatol=1e-5, maxiter=500):
    lib = _lib.load()
    seen = []

    def cb(x, _):
        seen.append(x)
        return float(fun(x))

    x, fx = C.c_double(), C.c_double()
    nfev, flag = C.c_int(), C.c_int()
    rc = lib.ree_scalar_brent(_lib.SCALAR_FN(cb), None, lo, hi, xatol, maxiter, C.byref(x), C.byref(fx), C.byref(nfev),
                              C.byref(flag))
    assert rc == 0
    return x.value, fx.value, nfev.value, flag.value, seen


def _hybr(fun, x0, xtol=1.49012e-08, maxfev=400, factor=100.0, epsfcn=np.finfo(float).eps):
    lib = _lib.load()
    seen = []

    def cb(x, _):
        seen.append(x)
        return float(fun(x))

    x, fx = C.c_double(), C.c_double()
    nfev, info = C.c_int(), C.c_int()
    rc = lib.ree_scalar_hybr1(_lib.SCALAR_FN(cb), None, x0, xtol, maxfev, factor, epsfcn, C.byref(x), C.byref(fx),
                              C.byref(nfev), C.byref(info))
    assert rc == 0
    return x.value, fx.value, nfev.value, info.value, seen


def _dedup(xs):
    out = []
    for x in xs:
        if not out or out[-1] != x:
            out.append(x)
    return out


# smooth, kinked, flat, multi-modal, infinite and NaN-producing objectives
BRENT_FUNS = [
    (lambda x: (x - 2.0) ** 2, 0.0, 5.0),
    (lambda x: (np.sqrt(1 + x * x) - 2.0) ** 2, 0.0, 0.7),
    (lambda x: abs(x - 1.234567) + 0.1 * x, 1.0, 3.0),
    (lambda x: np.cos(3 * x) + 0.1 * x, -2.0, 4.0),
    (lambda x: 1.0, 0.0, 1.0),
    (lambda x: x, 3.0, 3.0 + 1e-7),
    (lambda x: (np.tanh(5 * (x - 900.0)) - 0.3) ** 2, 871.3, 905.9),
    (lambda x: np.inf if x > 1.5 else (x - 1.4) ** 2, 0.0, 2.0),
    (lambda x: np.nan if x > 1.0 else (x - 0.9) ** 2, 0.0, 2.0),
    (lambda x: (np.exp(-x) * 50 - 2.0) ** 2, 0.0, 0.6931),
]


@pytest.mark.parametrize("case", range(len(BRENT_FUNS)))
def test_brent_matches_scipy_bit_for_bit(case):
    fun, lo, hi = BRENT_FUNS[case]
    ref_pts = []

    def rec(x):
        ref_pts.append(float(x))
        return fun(x)

    with np.errstate(all="ignore"):
        ref = minimize_scalar(rec, bounds=[lo, hi], method="Bounded")
        x, fx, nfev, flag, pts = _brent(fun, lo, hi)
    assert pts == ref_pts
    assert x == float(ref.x) or (np.isnan(x) and np.isnan(ref.x))
    assert nfev == ref.nfev and flag == ref.status


def test_brent_random_cvar_like_objectives():
    rng = np.random.default_rng(0)
    for _ in range(200):
        g = rng.standard_normal(400) * rng.uniform(0.2, 3) + rng.uniform(-1, 3)
        s0 = rng.uniform(0, 50)
        tgt = rng.uniform(0.5, 4)

        def ind(s):
            return 0.5 * (1 - s * g / np.sqrt(1 + s * s * g * g))

        def fun(s):
            w = ind(s) / ind(s0)
            return (np.std(w) / np.mean(w) - tgt) ** 2

        lo, hi = s0, s0 + rng.uniform(0.01, 20)
        ref_pts = []

        def rec(x):
            ref_pts.append(float(x))
            return fun(x)

        ref = minimize_scalar(rec, bounds=[lo, hi], method="Bounded")
        x, fx, nfev, flag, pts = _brent(fun, lo, hi)
        assert pts == ref_pts and x == float(ref.x) and nfev == ref.nfev


HYBR_FUNS = [
    (lambda x: x * x - 2.0, 1.0),
    (lambda x: np.exp(x) - 5.0, 0.0),
    (lambda x: np.tanh(x - 3.0), 0.5),
    (lambda x: x ** 3 - 2 * x - 5, 2.0),
    (lambda x: 1.0 / (1 + x * x) - 0.3, 0.1),
    (lambda x: 1.0 + x * x, 1.0),          # no root: slow progress exit
    (lambda x: 0.0 * x + 1.0, 2.0),        # zero jacobian
    (lambda x: x, 0.0),                    # starts at the root
    (lambda x: 1e6 * (x - 1e-3), 1.0),
    (lambda x: np.arctan(x - 40.0), 1.0),  # far start, flat tails
]


@pytest.mark.parametrize("case", range(len(HYBR_FUNS)))
def test_hybr_matches_scipy_bit_for_bit(case):
    fun, x0 = HYBR_FUNS[case]
    ref_pts = []

    def rec(x):  # the same Python-float arithmetic in both runs (numpy's array power differs from float ** by an ulp)
        ref_pts.append(float(np.asarray(x).reshape(-1)[0]))
        return fun(ref_pts[-1])

    with np.errstate(all="ignore"):
        ref = root(rec, x0)
        x, fx, nfev, info, pts = _hybr(fun, x0)
    # root() evaluates the start point once more before hybrd (shape check): compare without immediate repeats
    assert _dedup(pts) == _dedup(ref_pts)
    assert x == float(ref.x[0])
    assert info == ref.status
    assert nfev + 2 == ref.nfev  # scipy counts its two start-point probes (shape check, output allocation) as well


def test_hybr_random_ess_like_objectives():
    rng = np.random.default_rng(1)
    for _ in range(300):
        n = 500
        ell = -rng.gamma(2.0, rng.uniform(0.5, 30), n)
        tgt = rng.uniform(0.1, 0.9) * n

        def fun(b):
            b = float(np.asarray(b).reshape(-1)[0])
            a = b * ell
            w = np.exp(a - a.max())
            return w.sum() ** 2 / (w * w).sum() - tgt

        x0 = rng.uniform(0.01, 5)
        ref_pts = []

        def rec(x):
            ref_pts.append(float(np.asarray(x).reshape(-1)[0]))
            return fun(x)

        ref = root(rec, x0)
        x, fx, nfev, info, pts = _hybr(fun, x0)
        assert _dedup(pts) == _dedup(ref_pts)
        assert x == float(ref.x[0]) and info == ref.status and nfev + 2 == ref.nfev
