#!/usr/bin/env python
"""Benchmark of the CBREE hot path (BASELINE.json: particle-iterations/s and HBM fraction at N=1e7, d=100).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one CBREE iteration (solver.py:413-437 of the reference): step-size controller, sigma (Brent) and
beta (hybr) adaptation (one persistent search kernel), the Euler-Maruyama sweep and the Gram / Mahalanobis / IS
sweeps, convergence bookkeeping.  Workload: affine LSF, d=100, N=1e7 particles IN TOTAL (strong scaling: the ranks
share them; `weak_value` is the same step with 1e7 particles per GPU; --scaling weak makes that the headline), fp64,
synthetic standard-normal particles (Philox, rank-invariant counters).  Inputs (8 GB per ensemble, >= 1 GB per rank)
are far larger than the 126 MB L2, so no explicit L2 flush is needed between steps.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = the same metric through
CBREE(...).solve(problem) with the sample in pinned HOST memory (H2D of the sample and D2H of the results inside
the timed region); `roofline` = dominant sweep against the measured HBM peak (plus the fp64 pipe fraction that
actually binds at d=100); `cpu_baseline` = the reference's own CBREE (oracle/_ref; the NumPy oracle port if that copy
is absent) on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cbree_particle_iterations_per_sec"
UNIT = "particle-iterations/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=10_000_000,
                    help="particles: the TOTAL over all GPUs with --scaling strong, per GPU with --scaling weak")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): N total fixed at --n (BASELINE: 'near-linear 1->8 GPU scaling at N=10^7'); "
                         "weak: --n particles per GPU.  Identical at --gpus 1.")
    ap.add_argument("--no-weak", action="store_true", help="strong runs on >1 GPU: skip the extra weak-scaling leg")
    ap.add_argument("--d", type=int, default=100)
    ap.add_argument("--ref-n", type=int, default=100_000, help="particles of the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-fp64-probe", action="store_true", help="skip the in-line fp64 peak probe (profiling runs)")
    ap.add_argument("--no-code-warmup", action="store_true",
                    help="skip the small warm-up solve (profiling runs under ncu: its launches would fill the capture)")
    return ap.parse_args()


# --------------------------------------------------------------------------- #
# CPU arm: the reference's own CBREE (oracle/_ref, the git-ignored copy oracle/make_ref.sh makes; it ships with gpurun),
# else the NumPy oracle port of it.  The one place outside tests/ and smoke() that executes anything under oracle/.
# --------------------------------------------------------------------------- #
def cpu_cbree_reference(n, d, steps, warmup):
    """Time `steps` iterations of the UNMODIFIED reference `CBREE.solve` (solver.py:365-503) after `warmup` untimed ones,
    through its public API: the per-iteration `callback(cache, solver)` hook (solver.py:439-446) takes the time stamps.
    Returns (p-it/s, seconds) or None if no copy of the reference is reachable."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np

    import ref_loader

    if not ref_loader.reference_available():
        return None
    ref = ref_loader.load_reference()
    warmup = max(1, warmup)
    prob = ref.toy.make_linear_problem(d)
    prob.set_sample(np.random.default_rng(1).standard_normal((n, d)))  # (ERANataf.random takes 10 s per 1e8 values)
    stamps = []

    def callback(cache, solver):
        stamps.append(time.perf_counter())
        return cache

    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):  # the constructor prints "adaptive: True" (solver.py:267)
        solver = ref.solver.CBREE(seed=2, num_steps=warmup + steps - 1, convergence_check=False,
                                  divergence_check=False, callback=callback)
        sol = solver.solve(prob)
    if len(stamps) < warmup + steps:
        raise RuntimeError(f"reference CBREE stopped after {len(stamps)} iterations: {sol.msg}")
    dt = stamps[warmup + steps - 1] - stamps[warmup - 1]
    return n * steps / dt, dt


def cpu_cbree_port(n, d, steps, warmup):
    """Time `steps` CBREE iterations of the NumPy port after `warmup` untimed ones.  Returns (p-it/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np

    import cbree_oracle as co
    import lsf_oracle as lo

    rng = np.random.default_rng(1)
    X0 = rng.standard_normal((n, d))
    opts = co.Options(convergence_check=False, divergence_check=False, num_steps=10**9)
    orc = co.CbreeOracle(lo.lsf_affine, opts, rng=np.random.default_rng(2))
    st = co.IterState(X0, orc.lsf(X0), orc.e_fun(X0), sigma=opts.initial_sigma(), beta=opts.initial_beta())
    orc.compute_weights(st)
    orc.initial_stepsize(st)
    states = [st]

    def one():
        if len(states) > 1 and states[-1].iteration % 2 == 0:
            orc.update_stepsize(states)
        orc.update_beta_and_sigma(states[-1])
        states.append(orc.perform_step(states[-1]))
        if len(states) > opts.observation_window:
            states.pop(0)
        orc.convergence_check(states)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return n * steps / dt, dt


def cpu_cbree(n, d, steps, warmup):
    """(p-it/s, seconds, kind, description): the real reference when oracle/_ref (or /root/reference) is there."""
    use_all_host_threads()
    try:
        out = cpu_cbree_reference(n, d, steps, warmup)
    except Exception as e:  # say why the port ran instead
        out, why = None, f"; reference failed: {type(e).__name__}: {e}"
    else:
        why = "; oracle/_ref absent (run oracle/make_ref.sh where /root/reference exists)"
    if out is not None:
        return out[0], out[1], "reference", "rareeventestimation.solver.CBREE.solve (unmodified reference, oracle/_ref)"
    v, dt = cpu_cbree_port(n, d, steps, warmup)
    return v, dt, "port", "oracle/cbree_oracle.py (NumPy/SciPy port)" + why


def use_all_host_threads():
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm uses all the host threads it can get
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.pop(k, None)
    try:
        import numpy  # noqa: F401  (loads the BLAS the limits below apply to)
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def blas_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, dt, kind, what = cpu_cbree(args.ref_n, args.d, args.steps, args.warmup)
    cores = blas_threads()
    sample = (f"CBREE affine d={args.d}, N={args.ref_n} particles, {args.steps} iterations after {max(1, args.warmup)} "
              f"warm-up, {what} ({dt:.1f} s)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"CBREE affine LSF d={args.d} fp64 (bounded CPU sample N={args.ref_n} of BASELINE "
                               f"configs[1]; the rate is per particle-iteration)",
                   "particles": args.ref_n, "d": args.d, "host_cpus": os.cpu_count()},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- #
# clocks
# --------------------------------------------------------------------------- #
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- #
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import rareeventestimation_b200 as ree
    from rareeventestimation_b200 import _lib
    from rareeventestimation_b200.engine import Comm, Engine
    from rareeventestimation_b200.problem import DeviceSample, ShardedHostSample

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    eng = Engine.get(dev, Comm())
    d, K, W = args.d, args.steps, args.warmup
    strong = args.scaling == "strong"
    n_total = args.n if strong else args.n * world
    n_local = eng.comm.shard(n_total)[1] - eng.comm.shard(n_total)[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # per-call device timing of the two sweeps (CUDA events on the launching stream)
    marks = {"step": [], "moments": [], "maha": []}

    def timed(name, fn):
        def wrapper(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            marks[name].append((e0, e1))
            return out
        return wrapper

    eng.cbree_step = timed("step", eng.cbree_step)
    eng.cbree_moments = timed("moments", eng.cbree_moments)
    eng.cbree_maha = timed("maha", eng.cbree_maha)

    kw = dict(seed=1, convergence_check=False, divergence_check=False, quiet=True)
    # Code warm-up on a small ensemble of the same dimension: 16 iterations touch every kernel and both sigma-update
    # branches (the closed-form one of the first iterations and the Brent search that takes over around iteration 9).
    # CUDA loads kernels lazily at their first launch; inside the timed steps that was a sporadic +50 ms.
    if not args.no_code_warmup:
        warm = ree.make_linear_problem(d)
        warm.sample = DeviceSample(8192 * world, d, seed=1)
        ree.CBREE(num_steps=16, **kw).solve(warm)
        del warm
    # The sampler is started BEFORE the warm-up steps: launching nvidia-smi (NVML initialisation, device enumeration)
    # stalls CUDA calls of other processes for tens of milliseconds -- inside the timed region that was a sporadic
    # +20..100 ms.  Once it runs, its 200 ms polls are cheap; they cover the warm-up and the timed steps (same load).
    clocks = ClockSampler(local)
    if rank == 0 and os.environ.get("REE_BENCH_NO_CLOCKS") != "1":  # (hook used to measure the sampler's own cost)
        clocks.start()
        time.sleep(1.0)
    import gc

    def measure(n_tot):
        """W untimed + K timed CBREE iterations on n_tot particles (all ranks).  Returns a dict."""
        prob = ree.make_linear_problem(d)
        prob.sample = DeviceSample(n_tot, d, seed=0x5EED)
        solver = ree.CBREE(num_steps=10**9, **kw)
        cache_list = solver.start(prob, recycle=True)
        for _ in range(W):
            solver.advance(cache_list)
        for v in marks.values():
            v.clear()
        gc.collect()
        gc.disable()  # like timeit: no cyclic-GC pause inside the timed regions (re-enabled before the CPU baseline)
        barrier()
        launches0 = lib.ree_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(K):
            solver.advance(cache_list)
        ev1.record()
        barrier()
        gc.enable()
        out = {"launches": lib.ree_launch_count() - launches0, "ms": max_over_ranks(ev0.elapsed_time(ev1)),
               "sweep_ms": {k: (sum(a.elapsed_time(b) for a, b in v) / max(1, len(v))) for k, v in marks.items()},
               "sigma": cache_list[-1].sigma, "beta": cache_list[-1].beta}
        out["value"] = n_tot * K / (out["ms"] * 1e-3)
        return out  # the HBM blocks go back to the engine's free list / the allocator's cache

    weak = None
    if strong and world > 1 and not args.no_weak:  # second field: the same step with --n particles PER GPU
        weak = measure(args.n * world)
        eng.drop_pool()
        torch.cuda.empty_cache()
    m = measure(n_total)
    ms, launches, sweep_ms, value = m["ms"], m["launches"], m["sweep_ms"], m["value"]
    sigma_now, beta_now = m["sigma"], m["beta"]
    clk = clocks.stop() if rank == 0 else None
    gc.disable()

    # ---- roofline of the dominant sweep (per launch, this rank) -------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fp64_peak, fp64_src = None, None
    if rank == 0 and world == 1 and not args.no_fp64_probe and os.path.exists(os.path.join(ROOT, "tools", "fp64_peak")):
        # the fp64 roofline denominator measured in this very run (dependent-free DFMA / DMMA chains, < 1 s), with the
        # SM clock sampled while it runs
        try:
            probe_clk = ClockSampler(local)
            probe_clk.start()
            time.sleep(0.5)
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local)))
            txt = ""
            for _ in range(3):  # ~0.25 s each: long enough for the 200 ms clock samples to see the load
                txt = subprocess.run([os.path.join(ROOT, "tools", "fp64_peak")], capture_output=True, text=True,
                                     timeout=120, env=env).stdout
            fp64_peak = json.loads(txt.strip().splitlines()[-1])
            fp64_peak["clocks"] = probe_clk.stop()
            fp64_src = "tools/fp64_peak run inside this bench.py invocation"
        except Exception:
            fp64_peak = None
    if fp64_peak is None:
        try:
            fp64_peak = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))
            fp64_src = "profiles/fp64_peak.json (tools/fp64_peak.cu on this pool, an earlier run)"
        except Exception:
            pass
    # Algorithmic bytes / flops per particle of the three kernels of an iteration:
    #   step     read X, write X', g', e'                     | triangular F z: d(d+1) flop
    #   moments  read X', g', a                               | two symmetric-half Grams of (y, 1): 2 (d+1)(d+2) flop
    #   maha     read X', g', e'                              | triangular Linv y: d(d+1) flop
    # (d <= 55 runs the fused two-sweep path: "moments" is empty and the Grams are part of step / maha.)
    split = bool(marks["moments"])
    alg = {"step": n_local * (16 * d + 16), "moments": n_local * (8 * d + 16), "maha": n_local * (8 * d + 16)}
    flops = {"step": n_local * (d * (d + 1.0) + (0 if split else (d + 1.0) * (d + 2))),
             "moments": n_local * 2.0 * (d + 1) * (d + 2),
             "maha": n_local * (d * (d + 1.0) + (0 if split else (d + 1.0) * (d + 2)))}
    sweep_ms = {k: v for k, v in sweep_ms.items() if v > 0}
    dom = max(sweep_ms, key=lambda k: sweep_ms[k])
    names = {"step": "ree_cbree_step", "moments": "ree_cbree_moments", "maha": "ree_cbree_maha"}
    traffic = None
    try:
        tr_file = os.path.join(ROOT, "profiles", "r2_traffic.json")
        tr = json.load(open(tr_file if os.path.exists(tr_file) else os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if split and int(tr.get("d", -1)) == d:
            traffic = tr["bytes_per_particle"][names[dom]] * n_local
    except Exception:
        pass
    fp64_pk = max(fp64_peak.get("dfma_tflops", 0), fp64_peak.get("dmma_m8n8k4_tflops", 0)) if fp64_peak else None
    per_kernel = {}
    for k, ms_k in sweep_ms.items():
        gbs, tfl = alg[k] / (ms_k * 1e-3) / 1e9, flops[k] / (ms_k * 1e-3) / 1e12
        per_kernel[names[k]] = {"ms_per_launch": ms_k, "hbm_gbs": gbs, "hbm_frac": gbs / hbm_peak, "fp64_tflops": tfl,
                                "fp64_frac": tfl / fp64_pk if fp64_pk else None}
    ach_tf = flops[dom] / (sweep_ms[dom] * 1e-3) / 1e12
    ach_gb = alg[dom] / (sweep_ms[dom] * 1e-3) / 1e9
    iter_flops = sum(flops[k] for k in sweep_ms)
    if d > 35 and fp64_pk:
        # d = 100: 16.8 flop/B against a machine balance of 5.7 -> the fp64 tensor pipe (DMMA) binds, not HBM
        roofline = {"bound": "tensor", "kernel": names[dom], "achieved": ach_tf, "peak": fp64_pk, "unit": "TFLOP/s",
                    "frac": ach_tf / fp64_pk, "traffic": traffic,
                    "peak_source": f"fp64 DMMA (mma.sync.m8n8k4.f64) measured by tools/fp64_peak.cu: {fp64_src}; "
                                   "tcgen05 has no f64 kind, MEASURED_PEAKS.json only has bf16",
                    "fp64_probe": fp64_peak,
                    "ms_per_launch": sweep_ms[dom],
                    "hbm": {"achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gb / hbm_peak,
                            "peak_source": peak_src},
                    "kernels": per_kernel,
                    "whole_iteration": {"fp64_frac": iter_flops / (ms / K * 1e-3) / 1e12 / fp64_pk,
                                        "hbm_frac": (n_local * (24 * d + 48)) / (ms / K * 1e-3) / 1e9 / hbm_peak}}
    else:
        roofline = {"bound": "hbm", "kernel": names[dom], "achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach_gb / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "ms_per_launch": sweep_ms[dom], "kernels": per_kernel,
                    "whole_iteration": {"hbm_frac": (n_local * (24 * d + 48)) / (ms / K * 1e-3) / 1e9 / hbm_peak}}

    # ---- end to end through the reference-facing API with HOST buffers ---------------------------
    e2e = None
    if not args.no_e2e:
        lo, hi = eng.comm.shard(n_total)
        host = torch.empty((hi - lo, d), dtype=torch.float64, pin_memory=True)
        src = DeviceSample(n_total, d, seed=0x5EED).materialise(eng, lo, hi)
        eng.download(src, out=host.numpy())
        eng.release(src)  # back to the engine's free list: like a second solve in one process (do_multiple_solves),
        del src           # the timed call below reuses the HBM buffers of the solve above instead of cudaMalloc'ing
        prob2 = ree.make_linear_problem(d)
        prob2.sample = ShardedHostSample(host.numpy(), n_total, lo)
        solver2 = ree.CBREE(num_steps=K - 1, **kw)
        eng.reserve(hi - lo, d, 8)  # allocator warm-up (un-timed), the state of a process that has solved before
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        # solver2.solve(prob2), written out in its three public phases so that the line can say where the time goes
        t_ph = [time.perf_counter()]
        cl2 = solver2.start(prob2, recycle=True)                     # upload of the host sample, g / e, weights, initial step size
        torch.cuda.synchronize(); t_ph.append(time.perf_counter())
        while not cl2[-1].converged and cl2[-1].iteration <= solver2.num_steps:
            if not solver2.advance(cl2):
                break
        torch.cuda.synchronize(); t_ph.append(time.perf_counter())
        sol = solver2.finish(cl2)
        pf = float(sol.prob_fail_hist[-1])
        last_g = np.asarray(sol.lsf_eval_hist[-1])
        t_ph.append(time.perf_counter())
        f1.record()
        barrier()
        ms2 = max_over_ranks(f0.elapsed_time(f1))
        n_steps = sol.num_steps + 1  # trial step of the initial step size + num_steps iterations
        e2e = {"value": n_total * n_steps / (ms2 * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": (hi - lo) * d * 8 / n_steps,
               "d2h_bytes_per_step": (last_g.nbytes + 8 * len(sol.prob_fail_hist)) / n_steps,
               "steps": n_steps, "seconds": ms2 * 1e-3, "prob_fail": pf, "msg": sol.msg,
               "phases_ms": {"start_incl_h2d": 1e3 * (t_ph[1] - t_ph[0]), "iterations": 1e3 * (t_ph[2] - t_ph[1]),
                             "finish_incl_d2h": 1e3 * (t_ph[3] - t_ph[2])},
               "note": "CBREE.solve(problem) written out as its three public phases start / advance* / finish (the "
                       "body of solve) on a pinned HOST sample; D2H = lsf_eval_hist[-1] + prob_fail_hist.  The (S, N, d) "
                       "Solution.ensemble_hist is NOT copied: it stays in HBM behind LazyHistory until accessed."}
        del sol, solver2, host, cl2
        torch.cuda.empty_cache()

    gc.enable()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        steps_cpu = 8
        v, dt, kind, what = cpu_cbree(args.ref_n, d, steps_cpu, 1)
        cpu = {"value": v, "unit": UNIT, "cores": blas_threads(), "kind": kind,
               "sample": f"CBREE affine d={d}, N={args.ref_n}, {steps_cpu} iterations of {what} ({dt:.1f} s)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if (strong and world > 1) else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"CBREE affine LSF d={d}, N={n_total} particles in total, {n_local} per GPU, fp64 "
                                   f"(BASELINE configs[1])",
                       "particles_per_gpu": n_local, "particles_total": n_total, "d": d, "tgt_fun": "algebraic",
                       "noise": "philox", "sharding": f"particles x{world}",
                       "l2": "inputs (8*N*d bytes per ensemble) exceed L2; no flush",
                       "sigma_at_end": sigma_now, "beta_at_end": beta_now},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        if weak is not None:  # the same step at --n particles per GPU (N grows with the GPU count)
            line["weak_value"] = weak["value"]
            line["weak"] = {"value": weak["value"], "ms_per_step": weak["ms"] / K, "particles_per_gpu": args.n,
                            "particles_total": args.n * world, "gpu_launches": int(weak["launches"])}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
