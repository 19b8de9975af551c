"""Secondary paths on the engine's primitives: crude Monte Carlo (reference ``solver.py:1203-1271``) and sequential
importance sampling with adaptive conditional sampling MCMC (``solver.py:1064-1200`` -> ``sis/SIS_aCS.py:64-306``).
Same constructor arguments and ``Solution`` layout as the reference.
"""
from __future__ import annotations

import math
import warnings

import numpy as np
import torch
from scipy import optimize

from ._lib import check
from .engine import Engine, Ensemble, _pad
from .lsf import DeviceLSF
from .problem import DeviceSample, Problem, is_standard_normal_problem
from .solution import Solution
from .solver import Solver

_STREAM_CMC = 0x43 << 24
_STREAM_SIS = 0x49 << 24


class _LazyVector:
    """(1, N) history row that is downloaded on access."""

    def __init__(self, tensor, n):
        self._t, self.shape, self.ndim = tensor, (1, int(n)), 2

    def __array__(self, dtype=None, copy=None):
        return self._t[: self.shape[1]].cpu().numpy()[None, :]

    def __getitem__(self, idx):
        return np.asarray(self)[idx]


class _LazyEnsemble:
    """(1, N, d) history slice living in HBM."""

    def __init__(self, engine, ens):
        self._eng, self._ens, self.shape, self.ndim = engine, ens, (1, ens.n, ens.d), 3

    def __array__(self, dtype=None, copy=None):
        return self._eng.download(self._ens)[None, ...]

    def __getitem__(self, idx):
        return np.asarray(self)[idx]


class _LazySample:
    """(1, N, d) history of a counter-based sample: regenerated (Philox) and downloaded on access."""

    def __init__(self, sample: DeviceSample):
        self._s, self.shape, self.ndim = sample, (1,) + tuple(sample.shape), 3

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._s)[None, ...]

    def __getitem__(self, idx):
        return np.asarray(self)[idx]


def _device_lsf(prob: Problem, eng: Engine, d: int):
    if not isinstance(prob.lsf, DeviceLSF):
        return None
    return prob.lsf.descriptor(eng, d)


def _evaluate(prob: Problem, eng: Engine, ens: Ensemble):
    """g = lsf(ens) on the device, or through the host callback for user-supplied Python LSFs."""
    desc = _device_lsf(prob, eng, ens.d)
    if desc is not None:
        eng.lsf_eval(desc, ens, out=ens.g)
    else:
        g = np.asarray(prob.lsf(eng.download(ens)), dtype=np.float64).reshape(-1)
        ens.g[: ens.n].copy_(eng.to_device(g))


class CMC(Solver):
    """Crude Monte Carlo simulation (solver.py:1203-1271).

    noise="numpy" replays the reference exactly (batches of 1024 drawn on the host with per-batch seeds
    ``seed + i*num_runs`` through the problem's sampler); noise="philox" (default) draws the standard-normal sample
    on the device in chunks and only keeps the limit-state values -- the sample itself is a counter-based stream
    and is regenerated on access, so ``ensemble_hist`` (80 GB at 1e8 x 100) is never materialised."""

    def __init__(self, num_runs: int, seed: int = None, verbose=True, *, noise="philox", chunk=1 << 22) -> None:
        super().__init__()
        self.num_runs = int(num_runs)
        self.seed = seed
        self.verbose = verbose
        self.name = "Crude Monte Carlo Simulation"
        assert noise in ("philox", "numpy")
        self.noise = noise
        self.chunk = int(chunk)

    def __str__(self) -> str:
        return self.name

    def solve(self, prob: Problem) -> Solution:
        eng = Engine.get()
        d = prob.sample.shape[-1]
        if self.noise == "numpy" or not is_standard_normal_problem(prob):
            # the reference draws every batch through the problem's own sampler (solver.py:1248): the device stream
            # is N(0, I) and only stands in for it on the standard normal space
            return self._solve_reference_stream(prob, eng, d)
        n = self.num_runs
        lo, hi = eng.comm.shard(n)
        seed = 0 if self.seed is None else int(self.seed)
        g_all = eng.empty(_pad(hi - lo))
        fails, one = eng.zeros(1), eng.empty(1)
        for c0 in range(lo, hi, self.chunk):
            m = min(self.chunk, hi - c0)
            buf = eng.alloc_ensemble(m, d, first=c0)
            eng.philox_normal(m, d, seed, _STREAM_CMC, first=c0, out=buf)  # counter = global sample index
            _evaluate(prob, eng, buf)
            g_all[c0 - lo: c0 - lo + m].copy_(buf.g[:m])
            fails += eng.count_failed(buf, out=one, reduce=False)
            eng.release(buf)
        eng.comm.allreduce_sum_(fails)  # the only cross-rank exchange: one count
        pf = float(fails.item()) / n
        with np.errstate(all="ignore"):
            cvar = math.sqrt((1 - pf) / pf) if pf > 0 else math.nan  # scipy.stats.variation of the indicator
        sample = DeviceSample(n, d, seed, draw=_STREAM_CMC)
        prob.sample = sample  # the reference leaves the final sample in the problem (solver.py:1260)
        if self.verbose:
            print(f"{-(-n // 2**10)} batches. Current estimate: {pf}, cvar: {cvar}")
        return Solution(_LazySample(sample), math.nan * np.zeros(1), _LazyVector(g_all, hi - lo), np.ones(1) * pf, n,
                        "Success", other={"cvar": cvar})

    def _solve_reference_stream(self, prob, eng, d):
        batch_size = 2**10
        rest = self.num_runs % batch_size
        num_batches = self.num_runs // batch_size + (1 if rest != 0 else 0)
        lsf_evals = math.nan * np.zeros(self.num_runs)
        final_sample = math.nan * np.zeros((self.num_runs, d))
        msg, cost = "Success", 0
        for i in range(num_batches):
            batch_seed = None if self.seed is None else self.seed + i * self.num_runs
            bs = rest if (i == num_batches - 1 and rest != 0) else batch_size
            prob.set_sample(bs, seed=batch_seed)
            final_sample[cost:cost + bs, ...] = prob.sample
            cost += bs
        ens = eng.upload(final_sample)
        _evaluate(prob, eng, ens)
        lsf_evals[:] = ens.g[: ens.n].cpu().numpy()
        pf = np.sum(lsf_evals <= 0) / cost
        with np.errstate(all="ignore"):
            cvar = math.sqrt((1 - pf) / pf) if pf > 0 else math.nan
        if self.verbose:
            print(f"{num_batches} batches. Current estimate: {pf}, cvar: {cvar}")
        prob.sample = final_sample
        return Solution(final_sample[None, ...], math.nan * np.zeros(1), lsf_evals[None, ...], np.ones(1) * pf, cost, msg,
                        other={"cvar": cvar})


class SIS(Solver):
    """Sequential importance sampling; only the ``mixture_model="aCS"`` branch is on the engine's path
    (solver.py:1134-1147).  Keyword arguments as in the reference (solver.py:1070-1105)."""

    def __init__(self, **kwargs) -> None:
        super().__init__()
        self.num_chains_wrt_sample = kwargs.get("num_chains_wrt_sample", 0.1)
        self.burn_in = kwargs.get("burn_in", 0)
        self.cvar_tgt = kwargs.get("cvar_tgt", 1.0)
        self.mixture_model = kwargs.get("mixture_model", "GM")
        assert self.mixture_model in ["GM", "aCS", "vMFNM"], "mixture_model must be either 'GM' or 'aCS'"
        self.num_comps = kwargs.get("num_comps", 1)
        self.seed = kwargs.get("seed", None)
        self.fixed_initial_sample = kwargs.get("fixed_initial_sample", True)
        self.verbose = kwargs.get("verbose", False)
        self.return_other = kwargs.get("return_other", False)
        self.adapt_batches = kwargs.get("adapt_batches", 16)

    def __str__(self) -> str:
        return f"SiS ({self.mixture_model})"

    def solve(self, prob: Problem):
        if self.mixture_model != "aCS":
            raise NotImplementedError("only SIS with adaptive conditional sampling (mixture_model='aCS') is built; "
                                      "SIS-GM / SIS-vMFNM are out of scope (SURVEY.md section 2, rows 12-13)")
        assert hasattr(prob, "eranataf_dist"), \
            "For sequential importance sampling, prob needs to have an ERANataf distribution."
        if not is_standard_normal_problem(prob):
            raise NotImplementedError("SIS on the engine works on the standard normal space (NormalProblem); the U2X map "
                                      "of a general ERANataf distribution (transform_lsf, solver.py:1144) is host-side "
                                      "ERADist code that is out of scope (SURVEY.md section 2, row 10)")
        eng = Engine.get()
        comm = eng.comm
        lib = eng.lib
        N, d = prob.sample.shape
        p = self.num_chains_wrt_sample
        if (N * p != np.fix(N * p)) or (1 / p != np.fix(1 / p)):
            raise RuntimeError("N*p and 1/p must be positive integers. Adjust N and p accordingly")
        seed = int(np.random.SeedSequence(self.seed).generate_state(1)[0]) if self.seed is not None else \
            int(np.random.SeedSequence().generate_state(1)[0])
        seed = comm.broadcast_object(seed)  # one seed for all ranks (the unseeded case draws from OS entropy per process)
        nchain, lenchain, burn = int(N * p), int(round(1 / p)), int(self.burn_in)
        tar = float(self.cvar_tgt)
        max_it = 100
        # Multi-GPU: every rank owns a slice of the population; only sums cross ranks.  Resampling
        # moves no particle: the number of seeds a rank contributes is multinomial in the ranks' weight sums (drawn
        # from a generator all ranks share), each rank then resamples locally and runs its own chains, so the next
        # level's population of a rank is (its seeds) x (chain length) particles.
        lo, hi = comm.shard(N)
        # level-0 population (solver.py:1128-1131)
        if isinstance(prob.sample, DeviceSample):
            pop = prob.sample.materialise(eng, lo, hi)
        elif self.fixed_initial_sample:
            pop = eng.upload(np.asarray(prob.sample)[lo:hi], first=lo)
        else:
            pop = eng.philox_normal(hi - lo, d, seed, _STREAM_SIS, first=lo)
        _evaluate(prob, eng, pop)
        n_evals = N
        ws = eng.workspace(d)
        out4 = eng.empty(4)
        stream = lambda: eng.stream  # noqa: E731
        split_rng = np.random.default_rng(seed)  # identical on every rank

        def stats(ens, s_new, s_old, mode, w_out=None):
            """Global mean / std of the level weights (mode 0) or of the indicator (mode 1): plain sums, added over ranks."""
            if ens.n > 0:
                check(lib.ree_sis_reduce(ens.g.data_ptr(), ens.n, float(s_new), float(s_old), mode,
                                         w_out.data_ptr() if w_out is not None else None, ws.data_ptr(),
                                         out4.data_ptr(), stream()), "ree_sis_reduce")
            else:
                out4.zero_()
            t = comm.allreduce_sum_(out4).cpu().numpy()
            mean = t[1] / t[3]
            var = max(t[2] / t[3] - mean * mean, 0.0)
            return mean, math.sqrt(var), float(t[1])

        gsum = pop.g[:pop.n].sum().reshape(1)
        gmu = float(comm.allreduce_sum_(gsum).item()) / N
        sigmak = np.zeros(max_it + 1)
        Sk = np.ones(max_it)
        lam = 0.6
        cov_sl, m, draw = math.nan, 0, 0
        for m in range(max_it):
            s_old = sigmak[m] if m > 0 else -1.0

            def obj(x):  # |CoV(w(x)) - tarCoV|, SIS_aCS.py:153-178
                if not x > 0:
                    return math.inf
                mean, std, _ = stats(pop, x, s_old, 0)
                return abs(std / mean - tar) if mean > 0 else math.inf

            hi_s = 10.0 * gmu if m == 0 else sigmak[m]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                sigmak[m + 1] = optimize.fminbound(obj, 0, hi_s)
            w = eng.empty(max(pop.ld, 8))
            mean, _, _ = stats(pop, sigmak[m + 1], s_old, 0, w_out=w)
            Sk[m] = mean
            if Sk[m] == 0:
                break
            # how many of the nchain seeds each rank contributes: multinomial in the ranks' weight sums
            wsum = (w[:pop.n].sum() if pop.n > 0 else eng.zeros(1).sum()).reshape(1)
            if comm.enabled:
                pw = comm.all_gather_vec(wsum).cpu().numpy()
                counts = split_rng.multinomial(nchain, pw / pw.sum())
                my_chains, chain0 = int(counts[comm.rank]), int(counts[:comm.rank].sum())
            else:
                my_chains, chain0 = nchain, 0
            cur, cand = eng.alloc_ensemble(max(my_chains, 1), d), eng.alloc_ensemble(max(my_chains, 1), d)
            nxt = eng.alloc_ensemble(max(my_chains * lenchain, 1), d)
            cur.n = cand.n = my_chains
            nxt.n = my_chains * lenchain
            # seeds: multinomial resampling by scan + search + gather (SIS_aCS.py:194-196)
            draw += 1
            if my_chains > 0:
                cdf = eng.empty(pop.ld)
                scan_scratch = eng.empty(lib.ree_scan_scratch_doubles(pop.n))
                idx = torch.empty(my_chains, dtype=torch.int64, device=eng.device)
                check(lib.ree_scan_inclusive(w.data_ptr(), pop.n, cdf.data_ptr(), scan_scratch.data_ptr(), stream()), "scan")
                check(lib.ree_sis_resample(cdf.data_ptr(), pop.n, my_chains, seed, _STREAM_SIS + draw, chain0,
                                           idx.data_ptr(), stream()), "ree_sis_resample")
                check(lib.ree_gather_columns(pop.X.data_ptr(), d, pop.ld, idx.data_ptr(), my_chains, cur.X.data_ptr(),
                                             cur.ld, pop.g.data_ptr(), cur.g.data_ptr(), stream()), "ree_gather_columns")
            # aCS chains: all chains advance together; lambda is adapted between sub-batches of the chain steps
            # (the reference adapts after every 10 *sequential* chains, SIS_aCS.py:257-268 -- statistically
            # equivalent, not bit-reproducible)
            counta = 0
            alphas = []
            col = 0
            for step in range(lenchain + burn):
                sigmafk = min(lam, 1.0)
                rhok = math.sqrt(1 - sigmafk**2)
                draw += 1
                keep = step >= burn
                if my_chains > 0:
                    check(lib.ree_acs_propose(cur.X.data_ptr(), my_chains, d, cur.ld, rhok, sigmafk, seed,
                                              _STREAM_SIS + draw, chain0, cand.X.data_ptr(), stream()), "ree_acs_propose")
                    _evaluate(prob, eng, cand)
                    check(lib.ree_acs_accept(cur.X.data_ptr(), cur.g.data_ptr(), cand.X.data_ptr(), cand.g.data_ptr(),
                                             my_chains, d, cur.ld, float(sigmak[m + 1]), seed, _STREAM_SIS + draw, chain0,
                                             nxt.X.data_ptr() if keep else cand.X.data_ptr(),
                                             nxt.ld if keep else cand.ld, col if keep else 0,
                                             nxt.g.data_ptr() if keep else cand.g.data_ptr(), ws.data_ptr(),
                                             out4.data_ptr(), stream()), "ree_acs_accept")
                else:
                    out4.zero_()
                n_evals += nchain
                if keep:
                    col += my_chains
                t = comm.allreduce_sum_(out4).cpu().numpy()
                alpha_mu = t[1] / t[3]
                alphas.append(alpha_mu)
                counta += 1
                lam = min(math.exp(math.log(lam) + counta ** (-0.5) * (alpha_mu - 0.44)), 1.0)
            pop = nxt
            mean, std, _ = stats(pop, sigmak[m + 1], -1.0, 1)
            cov_sl = std / mean if mean > 0 else math.nan
            if self.verbose:
                print(f"\nCOV_Sl = {cov_sl}\n\t*aCS lambda = {lam}\t*aCS accrate = {np.mean(alphas)}")
            if cov_sl < tar:
                break
        l_tot = m + 1
        mean, _, _ = stats(pop, sigmak[m + 1], -1.0, 1)
        pr = mean * float(np.prod(Sk))
        other = {"k_fin": math.nan, "COV_Sl": cov_sl} if self.return_other else None
        return Solution(_LazyEnsemble(eng, pop), np.array([math.nan]), np.zeros((1, N)) * math.nan, np.array([pr]),
                        n_evals, "Success", num_steps=l_tot, other=other)
