"""Solvers with the reference's API (``src/rareeventestimation/solver.py``) on the B200 engine.

``CBREE.solve`` keeps the reference's *scalar* control flow on the host -- sigma/beta adaptation,
step-size controller, convergence checks, Solution assembly (solver.py:365-503) -- while every
N-sized operation runs in the CUDA kernels behind :class:`~.engine.Engine`.  The host only sees
d x d / scalar statistics.
"""
from __future__ import annotations

import logging
import math
import os
import warnings
from copy import deepcopy
from numbers import Real
from typing import Optional

import numpy as np
import torch
from scipy.optimize import minimize_scalar, root

from . import _lib
from .engine import DeviceLsf, Engine, Ensemble, cvar_from, ess_from
from .lsf import DeviceLSF
from .mixture import MixtureModel, VMFNMixture, vmfn_variates_numpy
from .problem import DeviceSample, Problem, ShardedHostSample
from .solution import Solution

_log = logging.getLogger(__name__)
_STREAM_STEP = 0x01 << 24  # Philox draw index base of the Euler-Maruyama noise
_STREAM_RESAMPLE = 0x02 << 24  # ... of the vMFNM resampling draws


def warn(msg):
    """The reference reports step failures through ``distutils.log.warn`` (solver.py:5, 430)."""
    warnings.warn(str(msg), RuntimeWarning, stacklevel=2)


class LazyHistory:
    """(S, N, d) or (S, N) history that lives in HBM and is downloaded on access."""

    def __init__(self, caches, attr):
        self._caches, self._attr = list(caches), attr
        first = getattr(caches[0], "_shape_" + attr)
        self.shape = (len(caches),) + tuple(first)
        self.ndim = len(self.shape)
        self.dtype = np.dtype(np.float64)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return getattr(self._caches[idx], self._attr)
        if isinstance(idx, tuple) and isinstance(idx[0], (int, np.integer)):
            return getattr(self._caches[idx[0]], self._attr)[idx[1:]]
        return np.asarray(self)[idx]

    def __array__(self, dtype=None, copy=None):
        return np.array([getattr(c, self._attr) for c in self._caches]).squeeze()


class CBREECache:
    """Per-iteration state with the reference's field names (solver.py:63-116).

    ``ensemble`` / ``lsf_evals`` / ``e_fun_evals`` are HBM resident; reading the attributes downloads
    them, assigning to them (as callbacks may, solver.py:439-446) uploads on the next use."""

    def __init__(self, engine: Engine, dev: Ensemble, sigma=1, beta=1, t_step=0.5, mixture_model=None,
                 num_lsf_evals=0):
        self._engine = engine
        self._dev = dev
        self._host = {}
        self._dirty = set()
        self._shape_ensemble = (dev.n, dev.d)
        self._shape_lsf_evals = (dev.n,)
        self._shape_e_fun_evals = (dev.n,)
        self._cw_dirty = False
        self.weighted_mean = None
        self.weighted_cov = None
        self.sigma = sigma
        self.beta = beta
        self.t_step = t_step
        self.cvar_is_weights = math.nan
        self.mixture_model = mixture_model if mixture_model is not None else MixtureModel(1)
        self.converged = False
        self.iteration = 0
        self.slope_cvar = math.nan
        self.sfp_slope = math.nan
        self.estimate = 0.0
        self.estimate_uniform_avg = 0.0
        self.estimate_sqrt_avg = 0.0
        self.estimate_cvar_avg = 0.0
        self.sfp = 0.0
        self.ess = math.nan
        self.num_lsf_evals = num_lsf_evals
        self.msg = ""
        # engine-side extras: unweighted moments of the ensemble and their device copies
        self.ens_mean = None
        self.ens_cov = None
        self.num_failed = 0.0
        self._stats = None  # device tensors of the last post-processing (see CBREE._post_ensemble)
        self._chol_info = 0

    FIELDS = ["ensemble", "lsf_evals", "e_fun_evals", "weighted_mean", "weighted_cov", "sigma", "beta", "t_step",
              "cvar_is_weights", "mixture_model", "converged", "iteration", "slope_cvar", "sfp_slope", "estimate",
              "estimate_uniform_avg", "estimate_sqrt_avg", "estimate_cvar_avg", "sfp", "ess", "num_lsf_evals", "msg"]

    # -- lazily materialised N-sized fields ---------------------------------- #
    def _get(self, name):
        if name not in self._host:
            eng, dev = self._engine, self._dev
            if name == "ensemble":
                self._host[name] = eng.download(dev)
            else:
                t = dev.g if name == "lsf_evals" else dev.e
                self._host[name] = eng.download_vector(t, dev.n)
        return self._host[name]

    def _set(self, name, value):
        self._host[name] = np.asarray(value, dtype=np.float64)
        self._dirty.add(name)

    ensemble = property(lambda s: s._get("ensemble"), lambda s, v: s._set("ensemble", v))
    lsf_evals = property(lambda s: s._get("lsf_evals"), lambda s, v: s._set("lsf_evals", v))
    e_fun_evals = property(lambda s: s._get("e_fun_evals"), lambda s, v: s._set("e_fun_evals", v))

    def sync_to_device(self) -> bool:
        """Upload fields a callback replaced.  Returns True if anything changed."""
        if not self._dirty:
            return False
        eng, dev = self._engine, self._dev
        if "ensemble" in self._dirty:
            new = eng.upload(self._host["ensemble"], first=dev.first)
            new.g, new.e = dev.g, dev.e
            if new.ld != dev.ld:
                new.g, new.e = eng.empty(new.ld), eng.empty(new.ld)
            self._dev = dev = new
            self._shape_ensemble = (dev.n, dev.d)
            self._shape_lsf_evals = (dev.n,)
        if "lsf_evals" in self._dirty:
            dev.g[: dev.n].copy_(eng.to_device(self._host["lsf_evals"]))
        if "e_fun_evals" in self._dirty:
            dev.e[: dev.n].copy_(eng.to_device(self._host["e_fun_evals"]))
        self._dirty.clear()
        return True

    def __deepcopy__(self, memo):
        new = CBREECache.__new__(CBREECache)
        for k, v in self.__dict__.items():
            if k in ("_engine", "_dev", "_stats"):
                new.__dict__[k] = v  # device buffers are shared copy-on-write (see own_device_buffers)
            else:
                new.__dict__[k] = deepcopy(v, memo)
        self._dev_shared = new._dev_shared = True
        return new

    def own_device_buffers(self) -> None:
        """Called before anything writes into ``_dev`` in place (the vMFNM redraw of the final importance sampling):
        a cache that shares its HBM buffers with a deep copy gets private ones first."""
        if getattr(self, "_dev_shared", False) and self._dev is not None:
            old = self._dev
            self._dev = Ensemble(old.X.clone(), old.n, old.d, old.ld, old.first,
                                 old.g.clone() if old.g is not None else None,
                                 old.e.clone() if old.e is not None else None)
            self._dev_shared = False


def flatten_cache_list(cache_list: list, attrs=None, lazy_bytes=64 << 20) -> dict:
    """dict attr -> stacked values over the caches (solver.py:119-136).  N-sized attributes above
    ``lazy_bytes`` stay in HBM behind a :class:`LazyHistory`."""
    if attrs is None:
        attrs = CBREECache.FIELDS
    out = dict.fromkeys(attrs)
    for a in attrs:
        if a in ("ensemble", "lsf_evals", "e_fun_evals"):
            shp = cache_list[0]._shape_ensemble if a == "ensemble" else cache_list[0]._shape_lsf_evals
            if len(cache_list) * int(np.prod(shp)) * 8 > lazy_bytes:
                out[a] = LazyHistory(cache_list, a)
                continue
        v = [getattr(c, a) for c in cache_list if getattr(c, a) is not None]
        out[a] = np.array(v).squeeze()
    return out


def get_slope(y) -> float:
    """Slope of a least-squares line through (i, y[i]); NaN if any entry is non-finite (utilities.py:145-163)."""
    y = np.atleast_1d(np.asarray(y, dtype=np.float64))
    if not np.all(np.isfinite(y)):
        return math.nan
    n = y.shape[0]
    if n < 2:  # polyfit of a single point fails (LAPACK error) and the reference returns nan for any failure
        return math.nan
    # closed form of polyfit(arange(n), y, 1)[0] (the reference's call, two 100 us SVD-based fits per iteration)
    x = np.arange(n, dtype=np.float64) - 0.5 * (n - 1)
    return float(np.dot(x, y - y.mean()) / np.dot(x, x))


class ConcatHistory:
    """History over all particles of a solve that was sharded over several devices of one process: the (S, N_r[, d])
    histories of the ranks side by side along the particle axis, materialised on access."""

    def __init__(self, parts):
        self._parts = list(parts)
        first = self._parts[0]
        n = sum(int(p.shape[1]) for p in self._parts)
        self.shape = (int(first.shape[0]), n) + tuple(first.shape[2:])
        self.ndim = len(self.shape)
        self.dtype = np.dtype(np.float64)

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        return np.concatenate([np.asarray(p).reshape((self.shape[0], -1) + self.shape[2:]) for p in self._parts], axis=1)

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return np.concatenate([np.asarray(p[idx]) for p in self._parts], axis=0)
        if isinstance(idx, tuple) and isinstance(idx[0], (int, np.integer)):
            return np.concatenate([np.asarray(p[idx[0]]) for p in self._parts], axis=0)[idx[1:]]
        return np.asarray(self)[idx]


def _merge_sharded_solutions(sols):
    """Rank 0's Solution (all scalar histories are identical on every rank) over the particles of all ranks."""
    out = sols[0]

    def two_d(h, want_ndim):  # a single retained cache comes back without its leading axis
        return h if h.ndim == want_ndim else np.asarray(h)[None, ...]

    out.ensemble_hist = ConcatHistory([two_d(s.ensemble_hist, 3) for s in sols])
    out.lsf_eval_hist = ConcatHistory([two_d(s.lsf_eval_hist, 2) for s in sols])
    out.N = out.ensemble_hist.shape[1]
    if isinstance(out.other, dict) and "cache_list" in out.other:
        out.other["cache_list_per_device"] = [s.other["cache_list"] for s in sols]
    return out


class Solver:
    """Methods to estimate the rare event probability of a `Problem` (solver.py:139-165)."""

    def __init__(self) -> None:
        pass

    def solve(self, prob: Problem):
        raise NotImplementedError

    # state a finished solve leaves behind that belongs to the engine / the problem, not to the solver's options:
    # shared by reference when the solver is copied (ctypes handles cannot be deep-copied; the next solve rebinds them)
    _SHARED_ON_COPY = ("_eng", "_prob", "_dev_lsf_obj", "_table", "lsf", "e_fun")

    def __deepcopy__(self, memo):
        """``set_options(..., in_situ=False)`` copies solvers *after* they have solved (the reference's
        do_multiple_solves does, convergence_analysis.py:70): every option, ``seed`` and ``rng`` are deep-copied,
        the engine handles of the last solve are not."""
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in vars(self).items():
            setattr(new, k, v if k in self._SHARED_ON_COPY else deepcopy(v, memo))
        return new

    def set_options(self, options: dict, in_situ=True):
        """Reset members already in ``vars(self)``; ``in_situ=False`` returns a configured deep copy."""
        if in_situ:
            for k, v in options.items():
                if k in vars(self):
                    setattr(self, k, v)
        else:
            other = deepcopy(self)
            other.set_options(options)
            return other


class _CountingLsf:
    """``solver.lsf``: counts evaluations like ``my_lsf`` in solver.py:376-381."""

    def __init__(self, fn):
        self.fn = fn
        self.counter = 0

    def __call__(self, x):
        self.counter += int(np.prod(np.shape(x)[:-1]))
        return self.fn(x)


class CBREE(Solver):
    """Consensus-based rare event estimation.  Constructor options and defaults are the reference's
    (solver.py:176-200); engine options are keyword-only additions that default to reference behaviour:

    noise: "philox" (device Philox4x32-10 counters, seeded by ``seed``) or "numpy" (draw Z on the host
           from ``self.rng`` exactly like ``rng.multivariate_normal`` does, then inject -- parity mode).
    search: "device" (the sigma / beta searches run inside one persistent kernel per iteration, ``ree_cbree_adapt``:
           SciPy's bounded Brent and MINPACK's hybrd in the same arithmetic) or "scipy" (the reference's host
           optimisers drive one device reduction per objective evaluation -- kept as the cross-check).
    devices: None (the current device; under torchrun one rank per GPU), an int N (devices 0..N-1) or a list of device
           ordinals: ``solve`` then shards the particles over these devices of THIS process (one host thread per
           device, NVLink mailboxes through peer access) and returns one Solution over all particles.
    """

    def __init__(self, stepsize_tolerance=0.5, t_step=-1, num_steps=100, tgt_fun="algebraic", observation_window=5,
                 seed=None, sigma_adaptivity="cvar", beta_adaptivity=True, cvar_tgt=2, sfp_tgt=1 / 3, ess_tgt=1 / 2,
                 lip_sigma=1, divergence_check=True, convergence_check=True, mixture_model="GM", resample=False,
                 verbose=False, save_history=False, return_other=False, return_caches=False, name=None,
                 callback=None, *, noise="philox", search=None, quiet=False, devices=None) -> None:
        super().__init__()
        if t_step > 0:
            self.stepsize_adaptivity = False
            self.t_step = t_step
        else:
            self.stepsize_adaptivity = True
        if not quiet:
            print(f"adaptive: {self.stepsize_adaptivity}")  # solver.py:267
        if isinstance(sigma_adaptivity, str):
            assert sigma_adaptivity in ["cvar", "sfp"]
            self.sigma = 0
            self.sigma_adaptivity = sigma_adaptivity
        else:
            assert sigma_adaptivity > 0, \
                f"sigma_adaptivity must be either `cvar`, `sfp` or `True`. I got '{sigma_adaptivity}'"
            self.sigma = sigma_adaptivity
            self.sigma_adaptivity = "constant"
        if beta_adaptivity is True:
            self.beta = 1
            self.beta_adaptivity = beta_adaptivity
        else:
            assert beta_adaptivity > 0, \
                f"beta_adaptivity must be either a positive number or `True`. I got '{beta_adaptivity}'"
            self.beta = beta_adaptivity
            self.beta_adaptivity = False
        self.num_comps = 1
        self.stepsize_tolerance = stepsize_tolerance
        self.num_steps = num_steps
        self.tgt_fun = tgt_fun
        self.observation_window = observation_window
        self.seed = seed
        self.rng = np.random.default_rng(self.seed)
        self.cvar_tgt = cvar_tgt
        self.sfp_tgt = sfp_tgt
        self.ess_tgt = ess_tgt
        self.lip_sigma = lip_sigma
        self.divergence_check = divergence_check
        self.convergence_check = convergence_check
        self.mixture_model = mixture_model
        self.resample = resample
        self.verbose = verbose
        self.save_history = save_history
        self.return_other = return_other
        self.return_caches = return_caches
        self.name = name
        self.callback = callback
        assert noise in ("philox", "numpy")
        self.noise = noise
        if search is None:
            search = os.environ.get("REE_SEARCH", "device")
        assert search in ("device", "scipy")
        self.search = search
        self.devices = devices

    def __str__(self) -> str:
        return self.name if self.name is not None else "CBREE"

    # ------------------------------------------------------------------ #
    # engine glue
    # ------------------------------------------------------------------ #
    def _tgt_kind(self) -> int:
        # unknown names fall through to the sigmoid like the reference's else-branch (solver.py:361)
        return _lib.TGT_KINDS.get(self.tgt_fun, _lib.TGT_KINDS["sigmoid"])

    def _setup(self, prob: Problem):
        if self.mixture_model not in ("GM", "vMFNM"):
            raise NotImplementedError(f"mixture_model={self.mixture_model!r}: CBREE knows 'GM' and 'vMFNM' (solver.py:186)")
        self._eng = Engine.get()
        self._prob = prob
        d_max = int(self._eng.lib.ree_max_dim())
        if prob.sample is not None and int(prob.sample.shape[-1]) > d_max:
            raise ValueError(f"rareeventestimation_b200 supports d <= {d_max}; the sample has d = "
                             f"{int(prob.sample.shape[-1])} (the d x d factorisations run in one CTA)")
        lsf = prob.lsf
        self._dev_lsf_obj = lsf if isinstance(lsf, DeviceLSF) else None
        self._device_energy = getattr(prob.e_fun, "is_standard_normal", False)
        self.lsf = _CountingLsf(prob.lsf)
        self.e_fun = prob.e_fun
        self._step_draws = 0
        self._philox_seed = int(self.rng.integers(0, 2**63 - 1)) if self.noise == "philox" else 0

    def _lsf_descriptor(self, d) -> Optional[DeviceLsf]:
        return self._dev_lsf_obj.descriptor(self._eng, d) if self._dev_lsf_obj is not None else None

    def _evaluate_external(self, dev: Ensemble, need_g: bool, need_e: bool):
        """Host callbacks for user-supplied Python LSFs / energies (README example): a D2H/H2D bridge
        around *user* code, not an engine fallback."""
        eng = self._eng
        X = eng.download(dev)
        if need_g:
            g = np.asarray(self._prob.lsf(X), dtype=np.float64).reshape(-1)
            dev.g[: dev.n].copy_(eng.to_device(g))
        if need_e:
            e = np.asarray(self._prob.e_fun(X), dtype=np.float64).reshape(-1)
            dev.e[: dev.n].copy_(eng.to_device(e))

    def _n_global(self, dev: Ensemble) -> int:
        return self._n_total

    def _initial_ensemble(self, prob: Problem) -> Ensemble:
        eng = self._eng
        sample = prob.sample
        n_total = int(sample.shape[0])
        self._n_total = n_total
        lo, hi = eng.comm.shard(n_total)
        if isinstance(sample, DeviceSample):
            dev = sample.materialise(eng, lo, hi)
        elif isinstance(sample, ShardedHostSample):
            if sample.lo != lo or sample.local.shape[0] != hi - lo:
                raise ValueError("ShardedHostSample does not match this rank's particle range")
            dev = eng.upload(sample.local, first=lo)
        else:
            dev = eng.upload(np.asarray(sample)[lo:hi], first=lo)
        desc = self._lsf_descriptor(dev.d)
        if desc is not None:
            eng.lsf_eval(desc, dev, out=dev.g)
        if self._device_energy:
            eng.energy(dev, out=dev.e)
        if desc is None or not self._device_energy:
            self._evaluate_external(dev, desc is None, not self._device_energy)
        self.lsf.counter += n_total
        return dev

    # ------------------------------------------------------------------ #
    # post-processing of an ensemble: moments, IS density, weights  (device)
    # ------------------------------------------------------------------ #
    def _post_ensemble(self, cache: CBREECache, sums_ready=None,
                       shift_u=None, shift_w=None, model=None, defer=False):
        """Everything the reference derives from (X, g, e, sigma, beta) of one cache:
        unweighted mean/cov + factorisation (solver.py:670-674), ESS shift (my_softmax), Mahalanobis sweep with
        IS statistics (683-686, 843-846) and the weighted mean/cov (629-636).

        Split path: what the NEXT iteration's control flow needs -- both means / covariances, the ESS tuple, the failure
        count -- is read back BEFORE the factorisation and the Mahalanobis / IS sweep are enqueued, so the host works on
        it while the GPU runs those; their results (pivot statistics, IS tuple) arrive with a second small readback.
        ``defer=True`` returns after the first one; :meth:`_finish_post` completes the cache."""
        # (the pinned readback slots belong to the engine: an unfinished cache of ANY solver on it is completed first)
        self._finish_post(getattr(self._eng, "_open_post", None))
        eng, dev = self._eng, cache._dev
        d = dev.d
        kind = self._tgt_kind()
        slen = d + d * d + 2
        # packed result buffer: [mean d | cov d*d | m_w d | C_w d*d | stats 8 | is4 4 | ess4 4 | sums_u tail 2 | sums_w tail 2]
        res = eng.empty(2 * d + 2 * d * d + 20)
        mean, cov = res[0:d], res[d:d + d * d]
        m_w, C_w = res[d + d * d:2 * d + d * d], res[2 * d + d * d:2 * d + 2 * d * d]
        o = 2 * d + 2 * d * d
        stats, is4, ess4 = res[o:o + 8], res[o + 8:o + 12], res[o + 12:o + 16]
        tails = res[o + 16:o + 20]
        L, Linv = eng.empty(d * d), eng.empty(d * d)
        early = eng.path(d) == 2 and sums_ready is None
        if early:
            # split path: log-weights + softmax shift, both Grams in one pass, moments of both
            a_vec = eng.scratch("logw", dev.ld)
            eng.reduce_ess(dev, cache.sigma, cache.beta, kind, out=ess4, a_out=a_vec)
            sums_uw = eng.empty(2 * slen)
            sums_u, sums_w = eng.cbree_moments(dev, a_vec, ess4[0:1], shift_u, sums_uw)
            eng.moments_chol(sums_u, shift_u, d, 1, False, mean, cov)  # mean / cov only; factorised after the readback
            eng.moments_chol(sums_w, shift_u, d, 0, True, m_w, C_w)
        else:
            if sums_ready is None:
                sums_u = eng.empty(slen)
                eng.ensemble_sums(dev, shift_u, sums_u)
            else:
                sums_u = sums_ready
            eng.moments_chol(sums_u, shift_u, d, 1, False, mean, cov, L, Linv, stats)
            eng.reduce_ess(dev, cache.sigma, cache.beta, kind, out=ess4)
            sums_w = eng.empty(slen)
            eng.cbree_maha(dev, mean, Linv, stats[0:1], cache.sigma, cache.beta, kind, ess4[0:1], sums_w, is4)
            eng.moments_chol(sums_w, mean, d, 0, True, m_w, C_w)  # weighted sums are centred at the unweighted mean
            if model is not None and getattr(model, "_is4_dev", None) is not None:
                is4.copy_(model._is4_dev)
            elif model is not None:  # IS statistics of the vMFN density instead of the Gaussian fit's
                model.logpdf_device(eng, dev, is4=is4)
        tails[0:2].copy_(sums_u[d + d * d:])
        tails[2:4].copy_(sums_w[d + d * d:])
        # factor of the weighted covariance for the NEXT step's noise, on a side stream: it overlaps the readback below,
        # the host's bookkeeping and the sigma / beta search instead of sitting between the search and sweep 1
        # (chol(scale C) = sqrt(scale) chol(C); the scale is only known after the search)
        F0 = F0_event = None
        if self.noise == "philox":
            side, main = eng.side_stream(), torch.cuda.current_stream(eng.device)
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                F0, f0stats = eng.empty(d * d), eng.empty(8)
                eng.chol_scaled(C_w, d, 1.0, F0, f0stats)
                F0_event = torch.cuda.Event()
                F0_event.record(side)
            res.record_stream(side)
            F0.record_stream(main)
        if early:
            host, ev_a = eng.to_host_async(res, 0)
            # enqueued behind the readback: factorisation (+ inverse, eigenvalue brackets) and the Mahalanobis / IS sweep
            eng.moments_chol(sums_u, shift_u, d, 1, False, mean, cov, L, Linv, stats)
            if model is None:
                eng.cbree_maha(dev, mean, Linv, stats[0:1], cache.sigma, cache.beta, kind, None, None, is4)
            elif getattr(model, "_is4_dev", None) is not None:  # already computed with g and e of the redraw
                is4.copy_(model._is4_dev)
            else:  # the ensemble was drawn from the fitted vMFN model: its density replaces the Gaussian fit
                model.logpdf_device(eng, dev, is4=is4)
            host_b, ev_b = eng.to_host_async(res[o:o + 12], 1)
            ev_a.synchronize()
        else:
            host = eng.to_host(res)  # the one synchronising D2H of the iteration (pinned)
            host_b, ev_b = host[o:o + 12], None
        self._peer_checks = getattr(self, "_peer_checks", 0) + 1
        if self._peer_checks % 16 == 1:  # (a 4-byte synchronising copy of the time-out flag: not every iteration)
            eng.comm.check_peers()
        cache.ens_mean = host[0:d].copy()
        cache.ens_cov = host[d:d + d * d].reshape(d, d).copy()
        cache.weighted_mean = host[d + d * d:2 * d + d * d].copy()
        cache.weighted_cov = host[2 * d + d * d:o].reshape(d, d).copy()
        e4, tl = host[o + 12:o + 16], host[o + 16:o + 20]
        cache._stats = dict(C_w=C_w, F0=F0, F0_event=F0_event)  # views into `res`; the O(N) scratch is released here
        cache._vmfn_density = model is not None
        cache.num_failed = float(tl[0])
        cache._ess4 = e4.copy()
        cache._n_total = float(tl[1])
        cache._pending_b = (host_b, ev_b)
        eng._open_post = cache
        if e4[3] <= 0:  # my_softmax on an all-non-finite vector (utilities.py:121)
            eng._open_post = None
            raise ValueError("attempt to get argmax of an empty sequence")
        if not defer:
            self._finish_post(cache)
        return cache

    def _finish_post(self, cache: Optional[CBREECache]) -> None:
        """Second readback of :meth:`_post_ensemble`: pivot statistics of the Gaussian fit and the IS tuple."""
        if cache is None or getattr(cache, "_pending_b", None) is None:
            return
        host_b, ev_b = cache._pending_b
        cache._pending_b = None
        if getattr(cache._engine, "_open_post", None) is cache:
            cache._engine._open_post = None
        if ev_b is not None:
            ev_b.synchronize()
        st, i4 = host_b[0:8], host_b[8:12]
        cache._chol_info = int(st[1])
        cache._min_pivot, cache._max_diag = float(st[2]), float(st[3])
        cache._tr_inv, cache._norm_inf = float(st[4]), float(st[5])
        cache._is4 = np.array(i4, dtype=np.float64)

    def _finish_is_stats(self, cache: CBREECache):
        """cvar_is_weights (solver.py:683-686) and the IS estimate (843-846) from the sweep-2 sums."""
        self._check_density(cache)
        cache.cvar_is_weights = cvar_from(cache._is4)
        cache.estimate = self._estimate_of(cache, in_loop=True)

    def _estimate_of(self, cache: CBREECache, in_loop=False) -> float:
        if not in_loop and not getattr(self, "_device_energy", True) and cache._dev is not None:
            return self._estimate_standard_target(cache)  # (inside the loop the reference computes no estimate at all)
        R, s1 = float(cache._is4[0]), float(cache._is4[1])
        if s1 == 0.0 or not math.isfinite(R):
            return 0.0
        return math.exp(R) * s1 / cache._n_total

    def _estimate_standard_target(self, cache: CBREECache) -> float:
        """The final estimate always uses the standard normal density as target (``gaussian_logpdf``, solver.py:843),
        whatever ``prob.e_fun`` is; cvar_is_weights uses ``e_fun`` (solver.py:683-686).  With a user-supplied energy
        the two differ, and the estimate is redone with |x|^2/2 + d/2 log 2pi: density of the fit (or of the fitted
        model) per particle on the device, the mean of the IS weights on the host (user-energy problems are bridged
        through the host anyway)."""
        eng, dev = self._eng, cache._dev
        d, kind = dev.d, self._tgt_kind()
        logq = eng.empty(dev.ld)
        model = self._fitted_model(cache)
        if model is not None and hasattr(model, "logpdf_device"):
            model.logpdf_device(eng, dev, logq=logq)
            lq = eng.download_vector(logq, dev.n)
        elif model is not None:
            lq = np.asarray(model.logpdf(cache.ensemble), dtype=np.float64).reshape(-1)
        else:
            slen = d + d * d + 2
            sums_u, sums_w = eng.empty(slen), eng.empty(slen)
            eng.ensemble_sums(dev, None, sums_u)
            mean, cov, L, Linv, stats = eng.empty(d), eng.empty(d * d), eng.empty(d * d), eng.empty(d * d), eng.empty(8)
            eng.moments_chol(sums_u, None, d, 1, False, mean, cov, L, Linv, stats)
            is4, ess4 = eng.empty(4), eng.empty(4)
            if eng.path(d) == 2:
                eng.cbree_maha(dev, mean, Linv, stats[0:1], cache.sigma, cache.beta, kind, None, None, is4, logq=logq)
            else:
                eng.reduce_ess(dev, cache.sigma, cache.beta, kind, out=ess4)
                eng.cbree_maha(dev, mean, Linv, stats[0:1], cache.sigma, cache.beta, kind, ess4[0:1], sums_w, is4,
                               logq=logq)
            lq = eng.download_vector(logq, dev.n)
        e_std = eng.download_vector(eng.energy(dev), dev.n)
        g = eng.download_vector(dev.g, dev.n)
        with np.errstate(all="ignore"):
            part = np.sum(np.exp(-e_std - lq) * (g <= 0))  # utilities.py:43-45
        part = eng.comm.allreduce_sum_(eng.to_device(np.array([part]))).item() if eng.comm.enabled else float(part)
        return part / cache._n_total

    _PSD_MSG = "When `allow_singular is False`, the input matrix must be symmetric positive definite."

    def _check_density(self, cache: CBREECache):
        """SciPy's multivariate_normal refuses numerically singular covariances: its ``_PSD`` raises
        LinAlgError if an eigenvalue is <= eps = 1e6 * eps_f64 * max|lambda| and ValueError if one is < -eps.  The
        factorisation brackets both extreme eigenvalues (min pivot >= lambda_min >= 1 / trace(cov^-1), max diag <=
        lambda_max <= |cov|_inf): singular for sure if min pivot <= thr * max diag, regular for sure if
        1 / trace(cov^-1) > thr * |cov|_inf; only in between are the eigenvalues themselves computed (host, d x d)."""
        if getattr(cache, "_vmfn_density", False):
            return  # no Gaussian fit behind this cache's importance density
        thr = 1e6 * np.finfo(np.float64).eps
        if cache._chol_info == 0 and cache._min_pivot > thr * cache._max_diag:
            tr_inv = getattr(cache, "_tr_inv", 0.0)
            if tr_inv > 0.0 and math.isfinite(tr_inv) and 1.0 / tr_inv > thr * cache._norm_inf:
                return
        elif cache._chol_info == 0 and math.isfinite(cache._min_pivot) and cache._min_pivot > 0:
            raise np.linalg.LinAlgError(self._PSD_MSG)  # positive definite in fp64, singular by SciPy's threshold
        cov = np.asarray(cache.ens_cov, dtype=np.float64)
        if not np.isfinite(cov).all():
            raise ValueError("array must not contain infs or NaNs")  # scipy _PSD -> eigh(check_finite=True)
        s = np.linalg.eigvalsh(cov)
        eps = thr * np.max(np.abs(s))
        if np.min(s) < -eps:
            raise ValueError("The input matrix must be symmetric positive semidefinite.")
        if np.any(s <= eps):
            raise np.linalg.LinAlgError(self._PSD_MSG)

    def _compute_weights_check(self, cache: CBREECache):
        if not np.isfinite(cache.weighted_cov).all():  # solver.py:637-638
            raise ValueError("Weighted covariance contains non-finite elements.")

    # ------------------------------------------------------------------ #
    # sigma / beta (host optimisers, device objectives)
    # ------------------------------------------------------------------ #
    def _sfp_now(self, cache: CBREECache) -> float:
        return cache.num_failed / cache._n_total

    def _update_sigma_sfp(self, cache: CBREECache) -> None:
        """solver.py:564-580."""
        current = self._sfp_now(cache)
        delta_max = self.lip_sigma * np.log(1 / cache.t_step)
        if self.sfp_tgt > np.finfo(float).eps:
            delta = np.sqrt(np.maximum(0, self.sfp_tgt - current)) * delta_max / np.sqrt(self.sfp_tgt)
        else:
            delta = 0
        cache.sigma += delta

    def _update_sigma_cvar(self, cache: CBREECache) -> None:
        """solver.py:525-562: closed form while too few particles failed, else bounded Brent on the
        squared CVAR mismatch; each objective evaluation is one ``ree_reduce_cvar`` sweep over g."""
        eng, dev, kind = self._eng, cache._dev, self._tgt_kind()
        try:
            if self._sfp_now(cache) < self.sfp_tgt:
                self._update_sigma_sfp(cache)
            else:
                out = eng.empty(4)

                def obj_fun(sigma):
                    t4 = eng.fetch(eng.reduce_cvar(dev, sigma, cache.sigma, kind, out=out))
                    return (cvar_from(t4) - self.cvar_tgt) ** 2

                opt_sol = minimize_scalar(
                    obj_fun, bounds=[cache.sigma, cache.sigma + self.lip_sigma * np.log(1 / cache.t_step)],
                    method="Bounded")
                cache.sigma = opt_sol.x
        except Exception as e:
            msg = f"Failed to update sigma: {str(e)}"
            if self.verbose:
                cache.msg += msg
            else:
                _log.info(f"Iteration {cache.iteration}: {msg}")

    def _ess(self, cache: CBREECache, beta, out=None, ell=None, range2=None) -> float:
        """ESS of the weights exp(beta*l).  `ell` (the log-target vector for cache.sigma, from ree_logtgt) makes
        repeated evaluations for different beta cost one exp per particle."""
        dev = cache._dev
        if ell is not None:
            t4 = self._eng.fetch(self._eng.reduce_scaled(ell, dev.n, dev.d, beta, out=out, range2=range2))
        else:
            t4 = self._eng.fetch(self._eng.reduce_ess(dev, cache.sigma, beta, self._tgt_kind(), out=out))
        if t4[3] <= 0:
            raise ValueError("attempt to get argmax of an empty sequence")
        return ess_from(t4)

    def _update_beta(self, cache: CBREECache) -> None:
        """solver.py:582-612: MINPACK hybr root of ESS(beta) - ess_tgt*N; objective = ``ree_reduce_ess``."""
        out = self._eng.empty(4)
        n_tot = cache._n_total
        ell = self._ell_of(cache)
        # the shift max(beta * l) of every evaluation follows from the range of l on this rank (ranks merge their tuples)
        rng2 = self._eng.vector_range(ell, cache._dev.n, cache._dev.d)

        seen = {}  # hybr evaluates its starting point three times (fcn, fdjac1, first step): one sweep is enough

        def obj_fun(temperature):
            b = float(np.asarray(temperature).reshape(-1)[0])
            if b not in seen:
                seen[b] = self._ess(cache, b, out, ell, rng2) - self.ess_tgt * n_tot
            return seen[b]

        try:
            sol = root(obj_fun, cache.beta)
            new_beta = cache.beta if sol.x.item() <= 0 else sol.x.item()
            cache.beta = new_beta
            if new_beta in seen:  # ESS(beta) of the accepted temperature is already known (solver.py:523)
                cache._ess_known = (new_beta, cache.sigma, seen[new_beta] + self.ess_tgt * n_tot)
        except Exception as e:
            msg = f"Failed to update beta: {str(e)}"
            if self.verbose:
                cache.msg += msg
            else:
                _log.info(f"Iteration {cache.iteration}: {msg}")

    def _ell_of(self, cache: CBREECache):
        """Device vector l(g, e; cache.sigma), computed once per sigma (the reference recomputes it inside every
        objective evaluation, solver.py:592-594)."""
        dev = cache._dev
        return self._eng.logtgt(dev.g, dev.e, dev.n, cache.sigma, self._tgt_kind(),
                                out=self._eng.scratch("ell", dev.ld))

    def _sigma_sfp_value(self, cache: CBREECache, t_step) -> float:
        """sigma after the closed-form update of solver.py:564-580 for step size ``t_step`` (nothing is modified)."""
        delta_max = self.lip_sigma * np.log(1 / t_step)
        if self.sfp_tgt > np.finfo(float).eps:
            delta = np.sqrt(np.maximum(0, self.sfp_tgt - self._sfp_now(cache))) * delta_max / np.sqrt(self.sfp_tgt)
        else:
            delta = 0
        return cache.sigma + delta

    def _search_args(self, cache: CBREECache, t_step) -> tuple:
        """What the search kernel is launched with for this cache and step size -- a pure function of the cache's
        host-side state: (sigma the searches start from, upper bound of the Brent search or None, failure note or None)."""
        sigma, sigma_hi, note = cache.sigma, None, None
        if self.sigma_adaptivity == "cvar":
            try:
                if self._sfp_now(cache) < self.sfp_tgt:
                    sigma = self._sigma_sfp_value(cache, t_step)
                else:
                    sigma_hi = cache.sigma + self.lip_sigma * np.log(1 / t_step)
                    if not (math.isfinite(cache.sigma) and math.isfinite(sigma_hi)):
                        raise ValueError("Optimization bounds must be finite scalars.")
                    if cache.sigma > sigma_hi:
                        raise ValueError("The lower bound exceeds the upper bound.")
            except Exception as e:
                sigma, sigma_hi, note = cache.sigma, None, f"Failed to update sigma: {str(e)}"
        if self.sigma_adaptivity == "sfp":
            sigma = self._sigma_sfp_value(cache, t_step)
        return (float(sigma), None if sigma_hi is None else float(sigma_hi), note, float(cache.beta),
                bool(self.beta_adaptivity), float(self.ess_tgt * cache._n_total), self._tgt_kind(), float(self.cvar_tgt))

    def _launch_search(self, cache: CBREECache, args: tuple, overlap=None, wait=True):
        sigma, sigma_hi, _, beta, do_beta, ess_n, kind, cvar_tgt = args
        return self._eng.adapt(cache._dev, kind, sigma, sigma_hi, cvar_tgt, beta, do_beta, ess_n, overlap=overlap, wait=wait)

    def _speculate_search(self, cache_list: list):
        """Launch the NEXT iteration's sigma / beta search now (behind the Mahalanobis sweep already in the stream): its
        inputs -- g, e, the failure count, the step size the controller will choose -- are all known once the moments
        of the new ensemble are on the host.  Nothing of the cache is modified; the next :meth:`advance` recomputes the
        arguments and takes the result over if they are identical (else it searches again)."""
        cache = cache_list[-1]
        t_next = cache.t_step
        if self.stepsize_adaptivity and len(cache_list) > 1 and cache.iteration % 2 == 0:
            t_next = self._stepsize_value(cache_list)
        args = self._search_args(cache, t_next)
        self._launch_search(cache, args, wait=False)
        pend = dict(cache=cache, args=args, t_next=t_next, prep=None, token=self._eng.adapt_launches)
        try:  # the step's host-side preparation as well (it only depends on the step size), while the GPU is busy
            keep, cache.t_step = cache.t_step, t_next
            pend["prep"] = self._prepare_step(cache)
        except Exception:  # noqa: BLE001  (the next advance prepares again and reports the error in the usual place)
            pend["prep"] = None
        finally:
            cache.t_step = keep
        return pend

    def _update_beta_and_sigma_device(self, cache: CBREECache, overlap=None, pending=None) -> None:
        """solver.py:505-523 with both searches inside one kernel (``ree_cbree_adapt``): one launch and one readback
        per iteration instead of ~25 host-driven objective evaluations.  Outcomes, messages and failure paths are those
        of :meth:`_update_sigma_cvar` / :meth:`_update_beta`."""
        eng = self._eng
        args = self._search_args(cache, cache.t_step)
        sigma0, sigma_hi, note = args[0], args[1], args[2]
        if note is not None:
            self._note_failure(cache, note)
        cache.sigma = sigma0  # (the closed-form update, if it applied)
        if (pending is not None and pending["cache"] is cache and pending["args"] == args
                and pending["token"] == eng.adapt_launches):
            if overlap is not None:  # the search was launched at the end of the previous iteration
                overlap()
            out = eng.adapt_result()
        else:
            out = self._launch_search(cache, args, overlap=overlap)
        cache._search = (int(out[5]), int(out[6]), int(out[3]), int(out[4]))  # evaluations and exit codes
        if sigma_hi is not None:
            if out[3] == -1:  # my_log_cvar on a vector without finite entries (utilities.py:92)
                self._note_failure(cache, "Failed to update sigma: zero-size array to reduction operation maximum "
                                          "which has no identity")
            else:
                cache.sigma = float(out[0])
        if self.beta_adaptivity and out[4] == -1:  # my_softmax on a vector without finite entries (utilities.py:121)
            self._note_failure(cache, "Failed to update beta: attempt to get argmax of an empty sequence")
        elif self.beta_adaptivity:
            cache.beta = float(out[1])
        if not out[11] > 0:
            raise ValueError("attempt to get argmax of an empty sequence")
        cache.ess = float(out[2])

    def _note_failure(self, cache: CBREECache, msg: str) -> None:
        if self.verbose:
            cache.msg += msg
        else:
            _log.info(f"Iteration {cache.iteration}: {msg}")

    def _update_beta_and_sigma(self, cache: CBREECache, overlap=None, pending=None) -> None:
        """solver.py:505-523.  ``overlap`` is host work that may run while the device searches (device path only)."""
        if self.search == "device" and self._eng.peers_connected():
            return self._update_beta_and_sigma_device(cache, overlap, pending)
        if overlap is not None:
            overlap()
        if self.sigma_adaptivity == "cvar":
            self._update_sigma_cvar(cache)
        if self.sigma_adaptivity == "sfp":
            self._update_sigma_sfp(cache)
        if self.beta_adaptivity:
            self._update_beta(cache)
        known = getattr(cache, "_ess_known", None)
        if known is not None and known[0] == cache.beta and known[1] == cache.sigma:
            cache.ess = known[2]
        else:
            cache.ess = self._ess(cache, cache.beta)

    # ------------------------------------------------------------------ #
    # the step
    # ------------------------------------------------------------------ #
    def _prepare_step(self, cache: CBREECache) -> dict:
        """Everything of :meth:`_perform_step` that does not depend on the sigma / beta of this iteration (runs on the
        host while the search kernel works): noise mean, conditioning shifts of the one-pass Grams, the small upload,
        the target ensemble."""
        eng, src = self._eng, cache._dev
        n, d = src.n, src.d
        t = float(cache.t_step)
        m_noise = (1 - t) * cache.weighted_mean
        shift_u_h = t * cache.ens_mean + m_noise  # predicted unweighted mean; the weighted Gram is shifted by m_w
        small = eng.to_device_small(np.concatenate([m_noise, shift_u_h, cache.weighted_mean]))
        return dict(t=t, small=small, dst=eng.alloc_ensemble(n, d, first=src.first), desc=self._lsf_descriptor(d))

    def _perform_step(self, cache: CBREECache, prep: Optional[dict] = None, defer: bool = False) -> CBREECache:
        """solver.py:642-694 (GM branch).  sweep 1 (``ree_cbree_step``) + post-processing (sweep 2).  With ``defer`` the
        new cache comes back as soon as the moments are on the host; :meth:`_complete_step` finishes it."""
        eng, src = self._eng, cache._dev
        n, d = src.n, src.d
        if isinstance(prep, dict) and "error" in prep:
            raise prep["error"]
        if not prep or prep.get("t") != float(cache.t_step):
            prep = self._prepare_step(cache)
        t, small, dst, desc = prep["t"], prep["small"], prep["dst"], prep["desc"]
        scale = (1 - t**2) * (1 + cache.beta)
        split = eng.path(d) == 2  # transform only; the moments come from the one-pass Gram kernel afterwards
        sums = None if split else eng.empty(d + d * d + 2)
        m_noise_d, shift_u, shift_w = small[0:d], small[d:2 * d], small[2 * d:3 * d]
        resample = self.resample and self.mixture_model == "vMFNM"
        step_desc = None if resample else desc  # the transformed ensemble of the resampling branch is only fitted
        if self.noise == "numpy":
            # Generator.multivariate_normal(method="svd"): mean + Z @ (U sqrt(S))^T   (solver.py:667)
            c_noise = scale * cache.weighted_cov
            Z = self.rng.standard_normal((self._n_total, d))
            lo, hi = eng.comm.shard(self._n_total)
            zdev = eng.upload(Z[lo:hi], first=src.first)
            u, s, _ = np.linalg.svd(c_noise)
            F = eng.to_device(u * np.sqrt(s))
            eng.cbree_step(src, dst, t, m_noise_d, F, step_desc, shift_u, sums, z=zdev)
        else:
            F0 = cache._stats.get("F0") if (cache._stats and not cache._cw_dirty) else None
            if F0 is not None:  # factored on the side stream while the host did its bookkeeping
                torch.cuda.current_stream(eng.device).wait_event(cache._stats["F0_event"])
                F = F0 * math.sqrt(scale)
            else:
                F = eng.empty(d * d)
                fstats = eng.empty(8)
                eng.chol_scaled(cache._stats["C_w"] if not cache._cw_dirty else eng.to_device(cache.weighted_cov), d,
                                scale, F, fstats)
            self._step_draws += 1
            eng.cbree_step(src, dst, t, m_noise_d, F, step_desc, shift_u, sums, seed=self._philox_seed,
                           draw=_STREAM_STEP + self._step_draws)
        model = None
        if resample:
            # solver.py:654-663: fit a one-component vMFN model to the moved ensemble, redraw the ensemble from it
            model = self._vmfn_redraw(dst)
            sums, shift_u = None, shift_w
        elif desc is None or not self._device_energy:
            self._evaluate_external(dst, desc is None, not self._device_energy)
            if desc is None and sums is not None:
                sums[d + d * d: d + d * d + 1].copy_(eng.count_failed(dst))
        self.lsf.counter += self._n_total
        new = CBREECache(eng, dst, mixture_model=model if model is not None else MixtureModel(1))
        new.iteration = cache.iteration + 1
        new.converged = cache.converged
        new.sigma = cache.sigma
        new.beta = cache.beta
        new.t_step = cache.t_step
        new.num_lsf_evals = self.lsf.counter
        new._cw_dirty = False
        self._post_ensemble(new, sums_ready=sums, shift_u=shift_u, shift_w=shift_w, model=model, defer=defer)
        if not defer:
            self._complete_step(new)
        return new

    def _complete_step(self, new: CBREECache) -> None:
        self._finish_post(new)
        self._finish_is_stats(new)
        self._compute_weights_check(new)

    def _vmfn_redraw(self, dev: Ensemble) -> VMFNMixture:
        """``model = VMFNMixture(1); model.fit(X); X = model.sample(N, rng)`` on the device ensemble ``dev`` (in place)
        followed by the limit state and the energy of the new points (solver.py:659-663, 676-680)."""
        eng = self._eng
        model = VMFNMixture(1)
        model.fit_device(eng, dev)
        if self.noise == "numpy":  # replay the reference's draws (Generator.gamma / beta / uniform / normal)
            variates = vmfn_variates_numpy(self.rng, self._n_total, dev.d, *model._scalars())
            model.sample_device(eng, dev, 0, 0, variates)
        else:
            self._step_draws += 1
            model.sample_device(eng, dev, self._philox_seed, _STREAM_RESAMPLE + self._step_draws)
        desc = self._lsf_descriptor(dev.d)
        model._is4_dev = None
        if desc is not None and desc.kind == _lib.LSF_AFFINE and self._device_energy:
            # one pass over the fresh ensemble: limit state, energy, model density and the IS statistics
            model._is4_dev = eng.empty(4)
            model.logpdf_device(eng, dev, is4=model._is4_dev, fuse_lsf=desc)
            return model
        if desc is not None:
            eng.lsf_eval(desc, dev, out=dev.g)
        if self._device_energy:
            eng.energy(dev, out=dev.e)
        if desc is None or not self._device_energy:
            self._evaluate_external(dev, desc is None, not self._device_energy)
        return model

    # ------------------------------------------------------------------ #
    # stopping
    # ------------------------------------------------------------------ #
    def _convergence_check(self, cache_list: list) -> None:
        """solver.py:696-738 (the window's failure counts come from the device reductions)."""
        last = cache_list[-1]
        iteration = last.iteration
        window = cache_list[-self.observation_window:]
        last.sfp = last.num_failed / last._n_total
        last.slope_cvar = get_slope(np.array([c.cvar_is_weights for c in window]))
        last.sfp_slope = get_slope(np.array([c.num_failed for c in window]))
        if self.sigma_adaptivity == "cvar":
            last.converged = bool((last.cvar_is_weights <= self.cvar_tgt) and self.convergence_check)
            if last.converged:
                last.msg += "Converged with given `cvar_tgt`."
            if self.divergence_check and iteration >= self.observation_window:
                sfp_mean = np.average([c.sfp for c in window])
                last.converged = bool(last.converged or (last.slope_cvar > 0.0 and sfp_mean >= self.sfp_tgt))
                if last.converged:
                    last.msg += "Converged due to `divergence_check`."
        if self.sigma_adaptivity == "sfp":
            last.converged = bool((last.sfp >= self.sfp_tgt) and self.convergence_check)
            if last.converged:
                last.msg += "Converged with given `sfp_tgt`."

    # ------------------------------------------------------------------ #
    # step size
    # ------------------------------------------------------------------ #
    @staticmethod
    def _stack(m, c):
        return np.concatenate((m[None, ...], np.tril(c)))

    def _update_stepsize(self, cache_list: list) -> None:
        cache_list[-1].t_step = self._stepsize_value(cache_list)

    def _stepsize_value(self, cache_list: list):
        """Embedded two-stage error estimate on (mean, tril cov) of the last two ensembles (solver.py:740-781);
        the moments were produced by the sweeps, nothing N-sized is touched.  Evaluated on the mean and the lower
        triangle only: the zeros of the upper triangle of ``tril`` add nothing to the sum of squares, they only count
        in the average (which the reference takes over all (d + 1) d entries)."""
        ch1, ch2 = cache_list[-2], cache_list[-1]
        h = -np.log(ch1.t_step)
        phi = (np.exp(-2 * h) - 1) / (-2 * h)
        bm1 = phi - 2 * (phi - 1) / (-2 * h)
        bm2 = 2 * (phi - 1) / (-2 * h)
        phi = (np.exp(-4 * h) - 1) / (-2 * h)
        bc1 = phi - 2 * (phi - 1) / (-2 * h)
        bc2 = 2 * (phi - 1) / (-2 * h)
        d = ch2.ens_mean.shape[0]
        il = self._tril_cache.get(d) if hasattr(self, "_tril_cache") else None
        if il is None:
            self._tril_cache = {d: np.tril_indices(d)}
            il = self._tril_cache[d]
        x1 = np.concatenate((ch1.ens_mean, ch1.ens_cov[il]))
        x = np.concatenate((ch2.ens_mean, ch2.ens_cov[il]))
        b1 = np.concatenate((np.full(d, bm1), np.full(il[0].shape[0], bc1)))
        b2 = np.concatenate((np.full(d, bm2), np.full(il[0].shape[0], bc2)))
        x_hat = 2 * h * (b1 * x1 + b2 * x)
        delta = x - x_hat
        w = self.stepsize_tolerance + self.stepsize_tolerance * np.maximum(abs(x), abs(x_hat))
        err = np.sqrt(np.sum((delta / w) ** 2) / ((d + 1) * d))
        q = 0.9 * np.sqrt(1 / err)
        return np.maximum(np.exp(-100), np.minimum(np.exp(-1e-5), np.exp(-q * h)))

    def _compute_initial_stepsize(self, cache: CBREECache) -> None:
        """solver.py:783-817, including the trial Euler step and the ``(f0 - f2/w)`` parenthesisation."""
        m0, c0 = cache.ens_mean, cache.ens_cov
        x0 = self._stack(m0, c0)
        w = self.stepsize_tolerance + self.stepsize_tolerance * abs(x0)
        d0 = np.sqrt(np.average((x0 / w) ** 2))
        f0 = self._stack(-m0 + cache.weighted_mean, -2 * c0 + 2 * cache.weighted_cov)
        d1 = np.sqrt(np.average((f0 / w) ** 2))
        h0 = 0.01 * d0 / d1
        cache.t_step = np.exp(-h0)
        self._update_beta_and_sigma(cache)
        cache_new = self._perform_step(cache)
        m2, c2 = cache_new.ens_mean, cache_new.ens_cov
        f2 = self._stack(-m2 + cache_new.weighted_mean, -2 * c2 + 2 * cache_new.weighted_cov)
        d2 = np.sqrt(np.average((f0 - f2 / w) ** 2)) / h0
        h1 = np.sqrt(0.01 / np.maximum(d1, d2))
        cache.t_step = np.exp(-np.maximum(100 * h0, h1))
        cache.num_lsf_evals = cache_new.num_lsf_evals
        dev, cache_new._dev = cache_new._dev, None  # the trial ensemble is discarded (solver.py:804-817)
        self._eng.release(dev)

    # ------------------------------------------------------------------ #
    # final estimates
    # ------------------------------------------------------------------ #
    def _importance_sampling(self, cache: CBREECache) -> None:
        """solver.py:819-846 (GM branch).  The statistics were accumulated by sweep 2 when the cache was
        produced; they are recomputed only if a callback replaced N-sized fields."""
        if cache.sync_to_device():
            self._refresh(cache)
        model = self._fitted_model(cache)
        if model is not None and not getattr(cache, "_vmfn_density", False) and cache._dev is not None:
            # a callback installed a fitted model without touching the N-sized fields: its density replaces the fit
            self._model_is_stats(cache, model)
        if self.mixture_model == "vMFNM" and not cache.mixture_model.fitted and cache._dev is not None:
            # solver.py:828-836: fit the model to the cache's ensemble, replace the ensemble by a sample of the model,
            # re-evaluate the limit state there; the estimate uses the model's density
            cache.own_device_buffers()
            cache.mixture_model = self._vmfn_redraw(cache._dev)
            cache._host.pop("ensemble", None)
            cache._host.pop("lsf_evals", None)
            self.lsf.counter += self._n_total
            cache.num_lsf_evals = self.lsf.counter
            is4 = cache.mixture_model._is4_dev
            if is4 is None:
                is4 = self._eng.empty(4)
                cache.mixture_model.logpdf_device(self._eng, cache._dev, is4=is4)
            cache._is4 = self._eng.fetch(is4)
            cache._vmfn_density = True
        self._check_density(cache)
        cache.estimate = self._estimate_of(cache)

    def _fitted_model(self, cache: CBREECache):
        """The cache's mixture model if it is fitted (then its density is the importance density, solver.py:825-826:
        the `callback_vmfnm` pattern of dimension_study.py:12-25), else None."""
        model = cache.mixture_model
        return model if getattr(model, "fitted", False) else None

    def _model_is_stats(self, cache: CBREECache, model) -> None:
        """IS statistics of a cache under a fitted model's density: [max r, sum 1[g<=0] exp(r-max), sum (..)^2, #finite r],
        r = -e - log q (solver.py:683-686, 843-846).  Engine models evaluate on the device; any other object with the
        reference's ``logpdf(sample)`` interface is *user* code and goes through the host bridge like a Python LSF."""
        eng, dev = self._eng, cache._dev
        if hasattr(model, "logpdf_device"):
            is4 = eng.empty(4)
            model.logpdf_device(eng, dev, is4=is4)
            cache._is4 = eng.fetch(is4)
        else:
            r = -np.asarray(cache.e_fun_evals) - np.asarray(model.logpdf(cache.ensemble), dtype=np.float64).reshape(-1)
            fin = np.isfinite(r)
            mult = (np.asarray(cache.lsf_evals) <= 0)[fin].astype(np.float64)
            top = np.max(r[fin]) if fin.any() else -math.inf
            w = np.exp(r[fin] - top) * mult
            cache._is4 = np.array([top, w.sum(), (w * w).sum(), float(fin.sum())])
        cache._vmfn_density = True

    def _refresh(self, cache: CBREECache):
        sig, bet = cache.sigma, cache.beta
        model = self._fitted_model(cache)
        if model is not None:
            model._is4_dev = None  # (statistics of an earlier redraw do not describe the replaced ensemble)
        self._post_ensemble(cache, model=model if hasattr(model, "logpdf_device") else None)
        if model is not None and not hasattr(model, "logpdf_device"):
            self._model_is_stats(cache, model)
        cache.sigma, cache.beta = sig, bet

    def _compute_weighted_estimates(self, cache_list: list) -> None:
        """solver.py:848-864."""
        k = min(len(cache_list), self.observation_window)
        sqrt_w = np.sqrt(np.arange(k))
        with np.errstate(all="ignore"):
            cvar_w = 1 / np.array([c.cvar_is_weights**2 for c in cache_list[-k:]])
        cvar_w[~np.isfinite(cvar_w)] = 0.0
        pfs = [c.estimate for c in cache_list[-k:]]
        last = cache_list[-1]
        last.estimate_uniform_avg = np.average(pfs)
        last.estimate_cvar_avg = np.average(pfs, weights=cvar_w) if np.sum(cvar_w) > 0 else 0.0
        last.estimate_sqrt_avg = np.average(pfs, weights=sqrt_w)

    # ------------------------------------------------------------------ #
    # main loop
    # ------------------------------------------------------------------ #
    def start(self, prob: Problem, recycle: bool = False) -> list:
        """Everything ``solve`` does before its loop (solver.py:375-409): upload / evaluate the initial ensemble,
        initial weights, initial step size.  Returns the cache list.  ``recycle=True`` is the caller's promise not to
        keep references to caches that left the observation window: their HBM buffers are then reused by later steps
        (what ``solve`` does with its private list)."""
        self._owns_cache_list = bool(recycle)
        self._setup(prob)
        dev = self._initial_ensemble(prob)
        if not self.save_history:  # the window's HBM buffers, allocated once and recycled by _retire
            # window + the ensemble being produced + the trial step of the step-size controller: no cudaMalloc of an
            # 8 GB ensemble inside the iteration loop (it showed up as a sporadic +50 ms in one of the first steps)
            self._eng.reserve(dev.n, dev.d, int(min(self.observation_window, self.num_steps + 1)) + 2)
        cache = CBREECache(self._eng, dev, sigma=self.sigma, beta=self.beta, num_lsf_evals=self.lsf.counter)
        self._post_ensemble(cache)
        self._compute_weights_check(cache)
        if self.stepsize_adaptivity:
            self._compute_initial_stepsize(cache)
        else:
            cache.t_step = self.t_step
        self._table = self._make_table() if self.verbose else None
        self._msg = "Success"
        return [cache]

    def advance(self, cache_list: list) -> bool:
        """One pass of the loop body of ``solve`` (solver.py:413-467).  False if the step or the callback failed
        (the text is kept for ``Solution.msg``)."""
        pending, self._pending = getattr(self, "_pending", None), None
        if pending is not None and pending["cache"] is not cache_list[-1]:
            self._drop_pending(pending)
            pending = None
        if self.stepsize_adaptivity and len(cache_list) > 1 and cache_list[-1].iteration % 2 == 0:
            if pending is not None:  # computed (from the same two caches) when the search was launched ahead
                cache_list[-1].t_step = pending["t_next"]
            else:
                self._update_stepsize(cache_list)
        prep = {}

        def prepare():  # host-side preparation of the step, overlapped with the search kernel
            if pending is not None and pending["prep"] is not None:
                prep.update(pending["prep"])
                return
            try:
                prep.update(self._prepare_step(cache_list[-1]))
            except Exception as e:  # noqa: BLE001  (re-raised inside the step's own error handling below)
                prep["error"] = e

        self._update_beta_and_sigma(cache_list[-1], overlap=prepare, pending=pending)
        # The next search can be launched before this step's Mahalanobis sweep has finished when nothing can change the
        # new cache in between (no callback) and the search runs on the device.
        speculate = self.callback is None and self.search == "device" and self._eng.peers_connected()
        try:
            new = self._perform_step(cache_list[-1], prep, defer=speculate)
            if speculate:
                self._pending = self._speculate_search(cache_list + [new])
                self._complete_step(new)
            cache_list.append(new)
        except Exception as e:
            self._msg = str(e)
            self._pending = None
            if not self.verbose:
                warn(str(e))
            return False
        if not self.save_history and len(cache_list) > self.observation_window:
            self._retire(cache_list.pop(0))
        self._convergence_check(cache_list)
        if self.callback is not None:
            try:
                cache_list[-1] = self.callback(cache_list[-1], self)
                if cache_list[-1].sync_to_device():
                    self._refresh_after_callback(cache_list[-1])
            except Exception as e:
                self._msg = str(e)
                if not self.verbose:
                    warn(str(e))
                return False
        if self.verbose:
            self._print_row(self._table, cache_list)
        return True

    def _drop_pending(self, pending) -> None:
        if pending is not None and pending.get("prep"):
            self._eng.release(pending["prep"].get("dst"))
            pending["prep"] = None

    def _retire(self, cache: CBREECache):
        """A cache left the observation window.  Its HBM buffers are recycled only inside :meth:`solve` -- there the
        cache list never leaves the solver before it is pruned (no callback, no ``return_caches``); users of the
        ``start`` / ``advance`` API hold the list themselves, so nothing they may still reference is reused."""
        if getattr(self, "_owns_cache_list", False) and not self.callback and not self.return_caches:
            dev, cache._dev = cache._dev, None
            self._eng.release(dev)

    def finish(self, cache_list: list) -> Solution:
        """solver.py:468-503: exit message, importance sampling of the retained caches, Solution."""
        msg = self._msg
        self._drop_pending(getattr(self, "_pending", None))  # (a search launched ahead for an iteration that never came)
        self._pending = None
        self._finish_post(getattr(self._eng, "_open_post", None))
        if not cache_list[-1].converged and cache_list[-1].iteration > self.num_steps:
            msg = "Not Converged."
        return self._build_solution(cache_list, msg, None)  # costs after the final importance sampling (solver.py:497)

    def _device_list(self):
        dv = self.devices
        if dv is None:
            return None
        dv = list(range(int(dv))) if isinstance(dv, (int, np.integer)) else [int(x) for x in dv]
        return dv if len(dv) > 1 else None

    def _solve_on_devices(self, prob: Problem, devices) -> Solution:
        """One host thread per device, every thread runs the single-rank ``solve`` on its particle shard with a
        :class:`~.engine.ThreadComm` (the ranks of a torchrun launch, inside one process).  All threads execute the
        identical scalar control flow on identical reduced values; rank 0's Solution is returned with the N-sized
        histories replaced by views over all shards."""
        import threading

        import torch

        from .engine import ThreadComm, ThreadGroup

        if torch.cuda.device_count() < len(devices):
            raise ValueError(f"devices={devices}: this process sees {torch.cuda.device_count()} CUDA device(s)")
        tgroup = ThreadGroup(devices)
        results, errors = [None] * len(devices), [None] * len(devices)

        def work(rank):
            dev = torch.device("cuda", devices[rank])
            eng = None
            try:
                torch.cuda.set_device(dev)
                eng = Engine.get(dev, ThreadComm(tgroup, rank))
                mine = deepcopy(self)  # same options, same generator state: identical host decisions on every rank
                mine.devices = None
                mine.quiet = True
                mine.verbose = self.verbose and rank == 0
                results[rank] = CBREE.solve(mine, prob)
            except BaseException as e:  # noqa: BLE001  (reported by the calling thread)
                errors[rank] = e
                tgroup.barrier.abort()
            finally:
                if eng is not None:  # later single-device solves on this device must not meet a dead group
                    from .engine import Comm

                    eng.comm, eng._comm_given = Comm(enabled=False), False

        threads = [threading.Thread(target=work, args=(r,), name=f"ree-dev{devices[r]}") for r in range(len(devices))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        first = next((e for e in errors if e is not None and not isinstance(e, threading.BrokenBarrierError)), None)
        if first is None:
            first = next((e for e in errors if e is not None), None)
        if first is not None:
            raise first
        return _merge_sharded_solutions(results)

    def solve(self, prob: Problem) -> Solution:
        """Estimate the rare event of `prob` (solver.py:365-503)."""
        devices = self._device_list()
        if devices is not None:
            return self._solve_on_devices(prob, devices)
        try:  # the list below is private to this call: retired caches' buffers are recycled
            cache_list = self.start(prob, recycle=True)
            while not cache_list[-1].converged and cache_list[-1].iteration <= self.num_steps:
                if not self.advance(cache_list):
                    break
            return self.finish(cache_list)
        finally:
            self._owns_cache_list = False

    def _refresh_after_callback(self, cache: CBREECache):
        """A callback replaced the ensemble / evaluations: redo the moments the next step needs.  The
        reference keeps the *old* weighted mean/cov in that case (solver.py:441, 693), so do we."""
        m_w, C_w, cv = cache.weighted_mean, cache.weighted_cov, cache.cvar_is_weights
        model = self._fitted_model(cache)
        if model is not None:
            model._is4_dev = None  # (statistics of an earlier redraw do not describe the callback's ensemble)
        self._post_ensemble(cache, model=model if hasattr(model, "logpdf_device") else None)
        if model is not None and not hasattr(model, "logpdf_device"):
            self._model_is_stats(cache, model)
        cache.weighted_mean, cache.weighted_cov, cache.cvar_is_weights = m_w, C_w, cv
        cache._cw_dirty = True

    def _build_solution(self, cache_list, msg, costs) -> Solution:
        for c in cache_list:
            self._importance_sampling(c)
        self._compute_weighted_estimates(cache_list)
        last = cache_list[-1]
        other = {  # NumPy scalars like the reference's (callers use ``.tolist()``, convergence_analysis.py:222)
            "Average Estimate": np.float64(last.estimate_uniform_avg),
            "Root Weighted Average Estimate": np.float64(last.estimate_sqrt_avg),
            "VAR Weighted Average Estimate": np.float64(last.estimate_cvar_avg),
            "CVAR": np.float64(last.cvar_is_weights),
            "SFP": np.float64(last.sfp),
        }
        tmp = flatten_cache_list(cache_list)
        tmp["sigma"] = np.insert(tmp["sigma"], 0, 0)  # thesis notation, solver.py:486-488
        tmp["beta"] = np.insert(tmp["beta"], 0, np.nan)
        tmp["t_step"] = np.insert(tmp["t_step"], 0, np.nan)
        if self.return_other:
            other = other | tmp
        if self.return_caches:
            other["cache_list"] = cache_list
        ens_hist = tmp["ensemble"]
        if not isinstance(ens_hist, LazyHistory) and ens_hist.ndim == 2:
            ens_hist = ens_hist[None, ...]
        costs = self.lsf.counter if costs is None else costs
        return Solution(ens_hist, tmp["beta"], tmp["lsf_evals"], tmp["estimate"], costs, msg,
                        num_steps=last.iteration, other=other)

    def solve_from_caches(self, cache_list: list) -> Solution:
        """Replay the stopping logic over precomputed caches without LSF calls (solver.py:866-971)."""
        self._eng = cache_list[0]._engine
        cache_list_new = [cache_list.pop(0)]
        table = self._make_table() if self.verbose else None
        msg = "Success"
        while (not cache_list_new[-1].converged and cache_list_new[-1].iteration <= self.num_steps
               and len(cache_list) > 0):
            cache_list_new.append(cache_list.pop(0))
            if not self.save_history and len(cache_list_new) > self.observation_window:
                cache_list_new.pop(0)
            self._convergence_check(cache_list_new)
            if self.callback is not None:
                try:
                    cache_list_new[-1] = self.callback(cache_list_new[-1], self)
                except Exception as e:
                    msg = str(e)
                    if not self.verbose:
                        warn(str(e))
                    break
            if self.verbose:
                self._print_row(table, cache_list_new)
        if not cache_list_new[-1].converged and cache_list_new[-1].iteration > self.num_steps:
            msg = "Not Converged."
        return self._build_solution(cache_list_new, msg, cache_list_new[-1].num_lsf_evals)

    # -- verbose table (solver.py:401-406, 448-467) ---------------------------- #
    def _make_table(self):
        cols = ["Iteration", "Sigma", "Beta", "Stepsize", "CVAR", "SFP", "Comment"]
        try:
            from prettytable import PrettyTable

            width = max(len(s) for s in cols)
            return PrettyTable(cols, float_format=".5", max_width=width, min_width=width)
        except ImportError:
            print(" | ".join(f"{c:>10s}" for c in cols))
            return None

    def _print_row(self, table, cache_list):
        c = cache_list[-1]
        row = [c.iteration, c.sigma, c.beta, np.log(1 / c.t_step), c.cvar_is_weights, c.sfp, cache_list[-2].msg]
        if table is None:
            print(" | ".join(f"{v:>10.5g}" if isinstance(v, Real) else f"{v!s:>10s}" for v in row))
        else:
            table.add_row(row)
            print(table.get_string(start=len(table.rows) - 1, end=len(table.rows), header=c.iteration == 1,
                                   border=False))
