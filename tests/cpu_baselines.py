"""CPU timings of the oracle (the NumPy / SciPy restatement of the reference) for the BASELINE.json configurations other
than the headline one, on bounded samples (PDE limit state at N=1e3, CMC at 1e5..1e6, SIS-aCS at N=1e4,
CBREE oscillator); one JSON line per configuration.  Test / measurement infrastructure: imports oracle/.
Usage: python tests/cpu_baselines.py > gpurun_out/cpu_baselines.jsonl"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import baselines_oracle as bo  # noqa: E402
import cbree_oracle as co  # noqa: E402
import lsf_oracle as lo  # noqa: E402


def threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def emit(**kw):
    kw["cores"] = threads()
    print(json.dumps(kw), flush=True)


# PDE limit state (diffusion.py:131-171): sparse assembly + SuperLU per particle
dif = lo.DiffusionOracle()
X = np.random.default_rng(0).standard_normal((1000, 150))
t0 = time.perf_counter()
g = dif.g(X)
dt = time.perf_counter() - t0
emit(config="PDE limit state d=150", sample="N=1000 evaluations of oracle/lsf_oracle.py DiffusionOracle", seconds=dt,
     lsf_evals_per_s=1000 / dt, note="vectorised Thomas restatement; the reference loops over particles with a sparse SuperLU solve (1.28 ms per particle, SURVEY.md 8a17)")

# CBREE on the oscillator (d=6)
n = 200_000
rng = np.random.default_rng(1)
X0 = rng.standard_normal((n, 6))
opts = co.Options(convergence_check=False, divergence_check=False, num_steps=10**9)
orc = co.CbreeOracle(lo.lsf_oscillator, opts, rng=np.random.default_rng(2))
st = co.IterState(X0, orc.lsf(X0), orc.e_fun(X0), sigma=opts.initial_sigma(), beta=opts.initial_beta())
orc.compute_weights(st)
orc.initial_stepsize(st)
states = [st]
t0 = time.perf_counter()
steps = 5
for _ in range(steps):
    if len(states) > 1 and states[-1].iteration % 2 == 0:
        orc.update_stepsize(states)
    orc.update_beta_and_sigma(states[-1])
    states.append(orc.perform_step(states[-1]))
    if len(states) > opts.observation_window:
        states.pop(0)
    orc.convergence_check(states)
dt = time.perf_counter() - t0
emit(config="CBREE oscillator d=6", sample=f"N={n}, {steps} iterations of oracle/cbree_oracle.py", seconds=dt,
     p_it_per_s=n * steps / dt)

# CMC, affine d=100 (solver.py:1225-1271: batches of 1024)
n = 200_000
t0 = time.perf_counter()
res = bo.cmc(lo.lsf_affine, 100, n, seed=1)
dt = time.perf_counter() - t0
emit(config="CMC affine d=100", sample=f"N={n} samples, oracle/baselines_oracle.py cmc", seconds=dt, samples_per_s=n / dt)

# SIS-aCS, affine d=100 (sis/SIS_aCS.py:64-306: sequential chains in Python)
n = 10_000
u0 = np.random.default_rng(3).standard_normal((n, 100))
t0 = time.perf_counter()
res = bo.sis_acs(lo.lsf_affine, u0, seed=4)
dt = time.perf_counter() - t0
evals = int(res[4])  # (pf, levels, population, cov, lsf evaluations)
emit(config="SIS-aCS affine d=100", sample=f"N={n} per level, {int(res[1])} levels, oracle/baselines_oracle.py sis_acs",
     seconds=dt, lsf_evals=evals, lsf_evals_per_s=evals / dt, pf=float(res[0]))
