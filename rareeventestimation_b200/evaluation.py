"""Repeated-solve harness and its post-processing on the engine.

Mirrors the reference's ``evaluation/convergence_analysis.py`` (``do_multiple_solves`` 34-147,
``study_cbree_observation_window`` 150-333, ``load_data`` 336-393, ``add_evaluations`` 396-479, ``aggregate_df``
482-519): same arguments, same CSV columns in the same order, same seeding convention
(run i: ``solver.set_options({"seed": i, "rng": default_rng(i)}, in_situ=False)``).  It is host-side bookkeeping around
``Solver.solve`` / ``CBREE.solve_from_caches``; the solves themselves run on the GPU.  No plotting.
"""
from __future__ import annotations

import gc
from ast import literal_eval
from copy import deepcopy
from glob import glob
from numbers import Real
from os import path
from tempfile import NamedTemporaryFile

import numpy as np
import pandas as pd
from numpy.random import default_rng
from scipy.stats import variation

from .solution import Solution
from .solver import CBREECache

_GROUP = ["Problem", "Solver", "Sample Size"]


def _new_csv(dir, prefix) -> str:
    with NamedTemporaryFile(prefix=prefix, suffix=".csv", delete=False, dir=dir) as f:
        return f.name


def _failed_solution(prob, exc) -> Solution:
    """What the reference records when ``solve`` itself raises (convergence_analysis.py:76-85)."""
    n = prob.sample.shape[0]
    return Solution(np.asarray(prob.sample)[None, ...], np.full(1, np.nan), np.full(n, np.nan), np.zeros(1), 0, str(exc))


def _cell(val):
    """One CSV cell for an entry of ``Solution.other``: arrays become lists, scalars stay."""
    if isinstance(val, np.ndarray):
        return val.tolist()
    if hasattr(val, "tolist") and not isinstance(val, (str, bytes)):
        try:
            return np.asarray(val).tolist()
        except Exception:
            return val
    return val


def _row(solver_name, prob, seed, solution, save_other, other_list, addtnl_cols) -> pd.DataFrame:
    df = pd.DataFrame(index=[0])
    df["Solver"] = solver_name
    df["Problem"] = prob.name
    df["Seed"] = seed
    df["Sample Size"] = prob.sample.shape[0]
    df["Truth"] = prob.prob_fail_true
    df["Estimate"] = solution.prob_fail_hist[-1]
    df["Cost"] = solution.costs
    df["Steps"] = solution.num_steps
    df["Message"] = solution.msg
    other = solution.other
    if other is not None:
        if save_other and other_list is None:
            for key in other.keys():
                df[key] = [_cell(other[key])]
        if other_list is not None:
            for key in other_list:
                val = _cell(other.get(key, np.asarray([pd.NA])))
                df[key] = [val]
                if not isinstance(val, (list, Real, str)):
                    print(f"{key} cannot be cast to a list and has type {type(val)}")
    if addtnl_cols is not None:
        for key, val in addtnl_cols.items():
            df[key] = val
    return df


def _report(estimates, i, num_runs, truth):
    rel = np.sqrt(np.average((estimates[: i + 1] - truth) ** 2)) / truth
    print(f"Rel. Root MSE after {i + 1}/{num_runs} runs: {rel}", end="\r" if i < num_runs - 1 else "\n")


def do_multiple_solves(prob, solver, num_runs: int, dir=".", prefix="", verbose=True, reset_dict=None,
                       save_other=False, other_list=None, addtnl_cols=None) -> str:
    """Call ``solver`` ``num_runs`` times on ``prob`` (seed i for run i) and append one CSV row per run.
    Returns the path of the CSV file (convergence_analysis.py:34-147)."""
    file_name = _new_csv(dir, prefix)
    estimates = np.zeros(num_runs)
    for i in range(num_runs):
        solver = solver.set_options({"seed": i, "rng": default_rng(i)}, in_situ=False)
        if reset_dict is not None:
            solver = solver.set_options(reset_dict, in_situ=False)
        try:
            solution = solver.solve(prob)
        except Exception as e:
            solution = _failed_solution(prob, e)
        df = _row(str(solver), prob, i, solution, save_other, other_list, addtnl_cols)
        df.to_csv(file_name, mode="a", header=(i == 0), index=False)
        if verbose:
            estimates[i] = solution.prob_fail_hist[-1]
            _report(estimates, i, num_runs, prob.prob_fail_true)
        del df, solution
        gc.collect()
    return file_name


def study_cbree_observation_window(prob, solver, num_runs: int, dir=".", prefix="", verbose=True,
                                   observation_window_range=range(2, 15), reset_dict=None,
                                   solve_from_caches_callbacks=None, save_other=False, other_list=None,
                                   addtnl_cols=None) -> str:
    """Per seed: one full CBREE run without the divergence check (history and caches kept), then a replay of the
    stopping logic (``solve_from_caches``) for every observation window and callback
    (convergence_analysis.py:150-333).  Rows carry the extra columns divergence_check, observation_window, callback;
    like the reference this file is written *with* the index column."""
    file_name = _new_csv(dir, prefix)
    write_header = True
    estimates = np.zeros(num_runs)
    original_callback = solver.callback
    for i in range(num_runs):
        solver = solver.set_options({"seed": i, "rng": default_rng(i), "divergence_check": False, "save_history": True,
                                     "return_caches": True, "callback": original_callback}, in_situ=False)
        if reset_dict is not None:
            solver = solver.set_options(reset_dict, in_situ=False)
        try:
            solution = solver.solve(prob)
            cache_list = solution.other["cache_list"]
            df = _row(solver.name, prob, i, solution, save_other, other_list, addtnl_cols)
            df["divergence_check"] = False
            df["observation_window"] = 0
            df["callback"] = solver.callback is not None
            df.to_csv(file_name, mode="a", header=write_header)
            write_header = False
            if solve_from_caches_callbacks is None:
                solve_from_caches_callbacks = {False: None}
            for win_len in observation_window_range:
                for callback_name, callback in solve_from_caches_callbacks.items():
                    solver = solver.set_options({"seed": i, "rng": default_rng(i), "divergence_check": True,
                                                 "observation_window": win_len}, in_situ=False)
                    if reset_dict is not None:
                        solver = solver.set_options(reset_dict, in_situ=False)
                    solver.callback = callback
                    try:
                        solution = solver.solve_from_caches(deepcopy(cache_list))
                    except Exception as e:
                        solution = _failed_solution(prob, e)
                    df = _row(solver.name, prob, i, solution, save_other, other_list, addtnl_cols)
                    df["divergence_check"] = True
                    df["observation_window"] = win_len
                    df["callback"] = callback_name
                    df.to_csv(file_name, mode="a", header=False)
            if verbose:
                estimates[i] = solution.prob_fail_hist[-1]
                _report(estimates, i, num_runs, prob.prob_fail_true)
        except Exception as e:  # the reference prints and moves on to the next seed
            print(e)
    return file_name


def load_data(pth: str, pattern: str, recursive=True, kwargs={}) -> pd.DataFrame:
    """All result files ``pth/pattern.csv`` as one frame; list-valued cache columns are parsed back into arrays
    (convergence_analysis.py:336-393, including its two repairs of historic files)."""
    frames = []
    for f in glob(path.join(pth, pattern + ".csv"), recursive=recursive):
        df = pd.read_csv(f, **kwargs)
        if {"divergence_check", "observation_window"} <= set(df.columns) and "True" in df["observation_window"].values:
            # files written with the two columns swapped
            df = df.rename(columns={"divergence_check": "observation_window", "observation_window": "divergence_check"})
            df = df.replace({"divergence_check": {"True": True, "0": False},
                             "observation_window": {"False": 0, "2": 2, "5": 5, "10": 10}})
        frames.append(df)
    df = pd.concat(frames, ignore_index=True)
    df = df.replace({"Problem": {"Fujita Rackwitz (d=2)": "Fujita Rackwitz Problem (d=2)",
                                 "Fujita Rackwitz (d=50)": "Fujita Rackwitz Problem (d=50)"}})
    df.drop_duplicates(inplace=True)
    df.reset_index(inplace=True)
    cache_fields = set(CBREECache.FIELDS)
    for c in df.columns:
        first = df[c][0] if len(df.index) else None
        if c in cache_fields and isinstance(first, str) and first.startswith("["):
            df[c] = df[c].apply(lambda cell: np.asarray(literal_eval(cell.replace("nan", "None"))))
    return df


def add_evaluations(df: pd.DataFrame, only_success=False, remove_outliers=False) -> pd.DataFrame:
    """Per (Problem, Solver, Sample Size) group: MSE family, quantiles of the relative error and of the cost, bias,
    success rate, coefficient of variation of the estimates -- written to every row of the group
    (convergence_analysis.py:396-479)."""
    df["Difference"] = df["Truth"] - df["Estimate"]
    df["Relative Error"] = abs(df["Difference"]) / df["Truth"]
    success = set(df.index[df["Message"] == "Success"])
    for _, idx in df.groupby(_GROUP).indices.items():
        idx = df.index[idx] if not isinstance(idx, pd.Index) else idx
        use = [i for i in idx if i in success] if only_success else list(idx)
        if remove_outliers:
            p75, p25 = df.loc[use, "Estimate"].quantile(q=[0.75, 0.25])
            use = [i for i in use if df.loc[i, "Estimate"] <= p75 + 3 * (p75 - p25)]
        est, diff, cost = df.loc[use, "Estimate"], df.loc[use, "Difference"], df.loc[use, "Cost"]
        rel = abs(diff) / df.loc[use, "Truth"]
        df.loc[idx, "Estimate Mean"] = np.average(est)
        df.loc[idx, "Estimate Variance"] = np.var(est)
        df.loc[idx, "Estimate Bias"] = df.loc[use, "Estimate Mean"] - df.loc[use, "Truth"]
        df.loc[idx, "MSE"] = np.average(diff ** 2)
        df.loc[idx, ".25 Relative Error"] = rel.quantile(q=0.25)
        df.loc[idx, ".75 Relative Error"] = rel.quantile(q=0.75)
        df.loc[idx, ".50 Relative Error"] = rel.quantile(q=0.5)
        df.loc[idx, "Relative MSE"] = df.loc[idx, "MSE"] / df.loc[idx, "Truth"] ** 2
        df.loc[idx, "Root MSE"] = np.sqrt(df.loc[idx, "MSE"])
        df.loc[idx, "Relative Root MSE"] = df.loc[idx, "Root MSE"] / df.loc[idx, "Truth"]
        df.loc[idx, "Relative Root MSE Variance"] = np.var(abs(df.loc[use, "Relative Error"]))
        df.loc[idx, "Cost Mean"] = np.average(cost)
        df.loc[idx, ".25 Cost"] = cost.quantile(q=0.25)
        df.loc[idx, ".75 Cost"] = cost.quantile(q=0.75)
        df.loc[idx, ".50 Cost"] = cost.quantile(q=0.5)
        df.loc[idx, "Cost Variance"] = np.var(cost)
        df.loc[idx, "Success Rate"] = np.average(df.loc[idx, "Message"] == "Success")
        df.loc[idx, "CVAR Estimate"] = variation(est)
    return df


def aggregate_df(df: pd.DataFrame, cols=None) -> pd.DataFrame:
    """One row per group (``cols`` + Problem, Solver, Sample Size): numbers are averaged, array cells are averaged
    element-wise, anything else keeps its first distinct value (convergence_analysis.py:482-519)."""
    cols = list(_GROUP) if cols is None else list(cols) + _GROUP
    value_cols = [c for c in df.columns if c not in _GROUP]

    def reduce_group(grp):
        out = []
        for c in value_cols:
            col = grp[c] if c in grp.columns else None
            if col is None or len(col) == 0:
                out.append(None)
            elif isinstance(col.values[0], np.ndarray):
                out.append(np.average(np.stack(col.values).astype(np.float64), axis=0))
            elif isinstance(col.values[0], Real):
                out.append(col.mean())
            else:
                out.append(col.unique()[0])
        return pd.Series(out, index=value_cols)

    rows = []
    for key, grp in df.groupby(cols):
        key = key if isinstance(key, tuple) else (key,)
        s = reduce_group(grp)
        for name, k in zip(cols, key):
            s[name] = k
        rows.append(s)
    out = pd.DataFrame(rows)
    lead = cols + [c for c in value_cols if c not in cols]
    return out[lead].reset_index(drop=True)
