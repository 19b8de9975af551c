"""CPU-only tests: the C-ABI library loads and exports what include/ree_b200.h declares, the host-side mirror
of the reference API behaves like the reference (options, set_options, problem protocol), the cross-rank
combination logic (world_size 2, gloo), and the loud failure without a GPU."""
import math
import os
import re
import sys

import numpy as np
import pytest
import torch

import lsf_oracle as lo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# --------------------------------------------------------------------------- #
# C ABI
# --------------------------------------------------------------------------- #
def header_symbols():
    text = open(os.path.join(ROOT, "include", "ree_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ree_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from rareeventestimation_b200 import _lib

    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/ree_b200.h but not exported"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) <= set(names)
    assert lib.ree_abi_version() == _lib.ABI_VERSION
    assert lib.ree_sums_len(100) == 100 + 100 * 100 + 2
    assert lib.ree_fused_path(100) == 2 and lib.ree_fused_path(6) == 1 and lib.ree_fused_path(150) == 2 and lib.ree_fused_path(200) == 0


def test_shared_object_is_sm100a_only():
    import subprocess

    from rareeventestimation_b200 import _lib

    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import rareeventestimation_b200 as ree
    from rareeventestimation_b200.engine import Engine

    with pytest.raises(ree.ReeError):
        Engine.get()
    prob = ree.make_linear_problem(4).set_sample(100, seed=1)
    with pytest.raises(ree.ReeError):
        ree.CBREE(seed=1, quiet=True).solve(prob)
    with pytest.raises(ree.ReeError):
        prob.lsf(prob.sample)  # LSFs are evaluated on the device too


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rareeventestimation_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("the oracle", ""), fn
            assert "/root/reference" not in src, fn


# --------------------------------------------------------------------------- #
# reference API surface
# --------------------------------------------------------------------------- #
def test_cbree_constructor_matches_reference_defaults(capsys):
    import rareeventestimation_b200 as ree

    s = ree.CBREE()
    assert "adaptive: True" in capsys.readouterr().out  # solver.py:267
    expect = dict(stepsize_tolerance=0.5, num_steps=100, tgt_fun="algebraic", observation_window=5, seed=None,
                  sigma_adaptivity="cvar", beta_adaptivity=True, cvar_tgt=2, sfp_tgt=1 / 3, ess_tgt=1 / 2,
                  lip_sigma=1, divergence_check=True, convergence_check=True, mixture_model="GM", resample=False,
                  verbose=False, save_history=False, return_other=False, return_caches=False, name=None,
                  callback=None, sigma=0, beta=1, num_comps=1, stepsize_adaptivity=True)
    for k, v in expect.items():
        assert getattr(s, k) == v, k
    assert str(s) == "CBREE" and str(ree.CBREE(name="x", quiet=True)) == "x"
    f = ree.CBREE(t_step=0.5, sigma_adaptivity=3.0, beta_adaptivity=2.0, quiet=True)
    assert (f.stepsize_adaptivity, f.t_step, f.sigma, f.sigma_adaptivity, f.beta, f.beta_adaptivity) == \
        (False, 0.5, 3.0, "constant", 2.0, False)
    with pytest.raises(AssertionError):
        ree.CBREE(sigma_adaptivity="nope", quiet=True)
    with pytest.raises(AssertionError):
        ree.CBREE(beta_adaptivity=-1.0, quiet=True)


def test_set_options_contract():
    """solver.py:151-165 as used by do_multiple_solves (convergence_analysis.py:70)."""
    import rareeventestimation_b200 as ree

    s = ree.CBREE(seed=1, quiet=True)
    s2 = s.set_options({"seed": 7, "rng": np.random.default_rng(7), "not_a_member": 3}, in_situ=False)
    assert s2 is not s and s2.seed == 7 and s.seed == 1 and not hasattr(s2, "not_a_member")
    assert s2.rng.standard_normal() == np.random.default_rng(7).standard_normal()
    assert s.set_options({"num_steps": 3}) is None and s.num_steps == 3


def test_problem_protocol_and_sample_parity():
    import rareeventestimation_b200 as ree

    p = ree.make_linear_problem(5)
    assert p.sample.shape == (1, 5) and p.name == "Linear Problem (d=5)"
    assert p.prob_fail_true == pytest.approx(2.3262907903552504e-4)
    assert p.set_sample(40, seed=3) is p
    np.testing.assert_array_equal(p.sample, lo.normal_sample(40, 5, 3))  # oracle is pinned to the reference
    X = np.ones((7, 5))
    assert p.set_sample(X).sample is X
    p.set_sample(np.ones((7, 4)))  # wrong dimension: message, sample unchanged (problem.py:72-73)
    assert p.sample is X
    assert hasattr(p, "eranataf_dist") and len(p.eranataf_dist.Marginals) == 5
    assert p.hints == {}
    dev = p.set_sample(1000, seed=1, device=True).sample
    assert dev.shape == (1000, 5)
    assert [q.name for q in ree.problems_lowdim] == ["Convex Problem", "Linear Problem (d=2)",
                                                     "Fujita Rackwitz Problem (d=2)"]
    assert ree.prob_nonlin_osc.prob_fail_true == pytest.approx(6.43e-6)
    assert len(ree.problems_highdim) == 18


def test_vectorizer_dispatch():
    import rareeventestimation_b200 as ree

    v = ree.Vectorizer(lambda x: float(np.sum(x)), 3)
    assert v(np.array([1.0, 2.0, 3.0])) == 6.0
    np.testing.assert_array_equal(v(np.arange(6.0).reshape(2, 3)), [3.0, 12.0])
    assert v.num_evals == 3
    assert ree.Vectorizer(v, 3).d == 3


def test_get_slope_and_tuple_formulas():
    import rareeventestimation_b200 as ree
    from rareeventestimation_b200.engine import cvar_from, ess_from

    u = np.load(os.path.join(ROOT, "tests", "golden", "unit_vectors.npz"))
    assert ree.get_slope([3.0, 2.5, 2.7, 1.0, 0.5]) == pytest.approx(float(u["slope"]), rel=1e-13)
    assert math.isnan(ree.get_slope([3.0, np.nan, 2.7]))
    # [M, S1, S2, cnt] -> ESS / cvar, against direct formulas
    a = np.random.default_rng(0).standard_normal(1000) * 3
    M = a.max()
    t4 = np.array([M, np.exp(a - M).sum(), np.exp(2 * (a - M)).sum(), len(a)])
    w = np.exp(a)
    assert ess_from(t4) == pytest.approx(w.sum() ** 2 / (w**2).sum(), rel=1e-12)
    assert cvar_from(t4) == pytest.approx(w.std() / w.mean(), rel=1e-10)
    assert cvar_from(np.array([0.0, 0.0, 0.0, 10.0])) == math.inf
    with pytest.raises(ValueError):
        cvar_from(np.array([-np.inf, 0.0, 0.0, 0.0]))


def test_diffusion_mode_table_matches_reference():
    import rareeventestimation_b200 as ree

    v = np.load(os.path.join(ROOT, "tests", "golden", "diffusion_vectors.npz"))
    lsf = ree.DiffusionLSF()
    tab = lsf.mode_table()
    assert tab.shape == (1536, 150)
    np.testing.assert_allclose(tab[:9], v["mode_table_head"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(tab.sum(0), v["mode_table_colsum"], rtol=1e-11, atol=1e-13)
    assert lsf.mu == pytest.approx(float(v["mu"]), rel=1e-15) and lsf.sigma == pytest.approx(float(v["sigma"]), rel=1e-15)


# --------------------------------------------------------------------------- #
# sharding: combination of per-rank partial statistics
# --------------------------------------------------------------------------- #
def local_expsums(a):
    fin = a[np.isfinite(a)]
    if fin.size == 0:
        return np.array([-np.inf, 0.0, 0.0, 0.0])
    M = fin.max()
    return np.array([M, np.exp(fin - M).sum(), np.exp(2 * (fin - M)).sum(), float(fin.size)])


def test_merge_expsums_equals_global_reduction():
    from rareeventestimation_b200.engine import merge_expsums

    a = np.random.default_rng(1).standard_normal(999) * 40 - 300
    a[5] = np.nan
    parts = np.stack([local_expsums(c) for c in np.array_split(a, 4)] + [local_expsums(np.array([np.nan]))])
    got = merge_expsums(torch.from_numpy(parts)).numpy()
    np.testing.assert_allclose(got, local_expsums(a), rtol=1e-13)
    empty = merge_expsums(torch.from_numpy(np.stack([local_expsums(np.array([np.inf]))] * 2))).numpy()
    assert empty[0] == -np.inf and empty[3] == 0


def _gloo_worker(rank, world, port, ret):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from rareeventestimation_b200.engine import Comm

    comm = Comm()
    n, d = 1001, 3
    rng = np.random.default_rng(42)
    X, a = rng.standard_normal((n, d)), rng.standard_normal(n) * 30
    lo_, hi_ = comm.shard(n)
    sums = torch.from_numpy(np.concatenate([X[lo_:hi_].sum(0), (X[lo_:hi_].T @ X[lo_:hi_]).ravel(), [hi_ - lo_]]))
    comm.allreduce_sum_(sums)
    t4 = comm.combine_expsums_(torch.from_numpy(local_expsums(a[lo_:hi_])))
    ret[rank] = (lo_, hi_, sums.numpy().copy(), t4.numpy().copy())
    dist.destroy_process_group()


def test_two_rank_combination_gloo():
    """world_size 2 on CPU (gloo): packed all_reduce of the raw sums and all_gather + merge of the max-shifted
    tuples give every rank the global statistics, bit-identical across ranks."""
    import torch.multiprocessing as mp

    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(world, port, ret), nprocs=world, join=True)
    n, d = 1001, 3
    rng = np.random.default_rng(42)
    X, a = rng.standard_normal((n, d)), rng.standard_normal(n) * 30
    assert (ret[0][0], ret[0][1], ret[1][0], ret[1][1]) == (0, 501, 501, 1001)
    want = np.concatenate([X.sum(0), (X.T @ X).ravel(), [n]])
    for r in range(world):
        np.testing.assert_allclose(ret[r][2], want, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(ret[r][3], local_expsums(a), rtol=1e-13)
    np.testing.assert_array_equal(ret[0][2], ret[1][2])
    np.testing.assert_array_equal(ret[0][3], ret[1][3])


def test_solver_is_deep_copyable_after_a_solve():
    """do_multiple_solves copies the solver of the previous run (convergence_analysis.py:70); the engine handles a
    solve leaves on the solver (ctypes / CUDA objects) must not break that, and options must not be shared."""
    import ctypes
    from copy import deepcopy

    import rareeventestimation_b200 as ree

    s = ree.CBREE(seed=3, quiet=True)
    s._eng = ctypes.pointer(ctypes.c_int(1))  # stands in for the engine: ctypes pointers refuse to be pickled / copied
    s._prob = object()
    s.lsf = lambda x: x
    c = s.set_options({"seed": 7, "rng": np.random.default_rng(7), "observation_window": 9}, in_situ=False)
    assert c is not s and c.seed == 7 and c.observation_window == 9 and s.seed == 3 and s.observation_window != 9
    assert c._eng is s._eng and c._prob is s._prob
    assert c.rng is not s.rng
    d = deepcopy(c)
    assert d.rng is not c.rng and d.rng.integers(0, 10**9) == deepcopy(c).rng.integers(0, 10**9)


def test_thread_comm_three_ranks_cpu():
    """The communicator of ``CBREE(devices=N)`` (one host thread per device, engine.ThreadComm) on CPU tensors with three
    threads: shards cover the particles, the packed sums are added in rank order (identical bits on every rank),
    objects are broadcast, vectors gathered in rank order; a failing rank releases the others instead of hanging them."""
    import threading

    import torch

    from rareeventestimation_b200.engine import ThreadComm, ThreadGroup

    world = 3
    tg = ThreadGroup(list(range(world)))
    n, d = 1000, 4
    X = np.random.default_rng(7).standard_normal((n, d)) * 1e3
    out = [None] * world

    def work(r):
        comm = ThreadComm(tg, r)
        lo, hi = comm.shard(n)
        part = torch.from_numpy(np.concatenate([X[lo:hi].sum(0), (X[lo:hi].T @ X[lo:hi]).ravel(), [hi - lo]]))
        total = comm.allreduce_sum_(part.clone())
        seed = comm.broadcast_object(1234 + r, src=1)
        gathered = comm.all_gather_vec(torch.tensor([float(lo), float(hi)], dtype=torch.float64))
        out[r] = (lo, hi, total.numpy().copy(), seed, gathered.numpy().copy())

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert [o[:2] for o in out] == [(0, 334), (334, 667), (667, 1000)]
    want = np.concatenate([X.sum(0), (X.T @ X).ravel(), [n]])
    for r in range(world):
        np.testing.assert_allclose(out[r][2], want, rtol=1e-12)
        np.testing.assert_array_equal(out[r][2], out[0][2])
        assert out[r][3] == 1235
        np.testing.assert_array_equal(out[r][4], [0, 334, 334, 667, 667, 1000])
    # a rank that dies aborts the barrier: the others raise instead of waiting forever
    tg2 = ThreadGroup([0, 1])
    errs = []

    def survivor():
        try:
            ThreadComm(tg2, 0).allreduce_sum_(torch.zeros(3, dtype=torch.float64))
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    t = threading.Thread(target=survivor)
    t.start()
    tg2.barrier.abort()
    t.join(timeout=30)
    assert not t.is_alive() and len(errs) == 1


def test_get_slope_matches_polyfit():
    """The closed form behind get_slope equals the reference's polyfit call (utilities.py:145-163) to rounding."""
    import warnings

    import rareeventestimation_b200 as ree

    rng = np.random.default_rng(3)
    for n in [2, 3, 5, 14, 40]:
        y = rng.standard_normal(n) * 10.0 ** int(rng.integers(-3, 4))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = np.polyfit(np.arange(n), y, deg=1, full=True)[0][0]
        assert ree.get_slope(y) == pytest.approx(want, rel=1e-12, abs=1e-15 * np.abs(y).max())
    assert math.isnan(ree.get_slope([1.0]))
