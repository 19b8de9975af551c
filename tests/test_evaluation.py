"""Repeated-solve harness (rareeventestimation_b200/evaluation.py) against the reference's
evaluation/convergence_analysis.py: post-processing functions compared with the live reference on the same CSV files
(build container), the harness itself run on the GPU like the reference's test/test_do_multiple_solves.py and
test/test_study_cbree_observation_window.py."""
import os
import tempfile

import numpy as np
import pandas as pd
import pytest


def _fake_results(dirname, seed=0):
    """Two result files as do_multiple_solves / study_cbree_observation_window write them."""
    rng = np.random.default_rng(seed)
    rows = []
    for prob, truth in [("Convex Problem", 4.21e-3), ("Linear Problem (d=2)", 2.33e-4)]:
        for solver in ["CBREE", "CBREE (vMFNM)"]:
            for n in [1000, 2000]:
                for s in range(7):
                    est = truth * np.exp(0.3 * rng.standard_normal())
                    rows.append({"Solver": solver, "Problem": prob, "Seed": s, "Sample Size": n, "Truth": truth,
                                 "Estimate": est, "Cost": int(n * rng.integers(5, 20)), "Steps": int(rng.integers(5, 20)),
                                 "Message": "Success" if rng.random() < 0.8 else "Not Converged.",
                                 "t_step": str([float(x) for x in np.round(rng.random(4), 6)]), "sigma": str([1.0, float("nan"), 3.5]),
                                 "col": "val"})
    df = pd.DataFrame(rows)
    df.iloc[: len(df) // 2].to_csv(os.path.join(dirname, "res_a.csv"), index=False)
    df.iloc[len(df) // 2:].to_csv(os.path.join(dirname, "res_b.csv"), index=False)
    return df


def _same_frame(a, b):
    assert list(a.columns) == list(b.columns)
    assert len(a.index) == len(b.index)
    for c in a.columns:
        x, y = a[c].values, b[c].values
        if len(x) and isinstance(x[0], np.ndarray):
            for u, v in zip(x, y):
                np.testing.assert_allclose(np.asarray(u, dtype=float), np.asarray(v, dtype=float), rtol=1e-13, equal_nan=True)
        elif a[c].dtype.kind in "fiu":
            np.testing.assert_allclose(x.astype(float), y.astype(float), rtol=1e-12, atol=0, equal_nan=True, err_msg=c)
        else:
            assert list(x) == list(y), c


@pytest.mark.needs_reference
@pytest.mark.parametrize("only_success,remove_outliers", [(False, False), (True, False), (False, True)])
def test_postprocessing_matches_reference(only_success, remove_outliers):
    import ref_loader

    from rareeventestimation_b200 import evaluation as ev

    ref = ref_loader.load_reference_evaluation()
    with tempfile.TemporaryDirectory() as d:
        _fake_results(d)
        a, b = ev.load_data(d, "res_*"), ref.load_data(d, "res_*")
        # glob order is the same for both; rows are identical frames
        _same_frame(a, b)
        assert isinstance(a["t_step"][0], np.ndarray) and a["sigma"][0][1] is None  # nan cells come back as None
        ea = ev.add_evaluations(a.copy(), only_success=only_success, remove_outliers=remove_outliers)
        eb = ref.add_evaluations(b.copy(), only_success=only_success, remove_outliers=remove_outliers)
        _same_frame(ea, eb)
        ga, gb = ev.aggregate_df(ea.drop(columns=["index"])), ref.aggregate_df(eb.drop(columns=["index"]))
        _same_frame(ga, gb)


def test_postprocessing_known_answers():
    from rareeventestimation_b200 import evaluation as ev

    df = pd.DataFrame({"Solver": ["S"] * 4, "Problem": ["P"] * 4, "Sample Size": [10] * 4, "Truth": [2.0] * 4,
                       "Estimate": [1.0, 2.0, 3.0, 6.0], "Cost": [10, 20, 30, 40],
                       "Message": ["Success", "Success", "Not Converged.", "Success"]})
    out = ev.add_evaluations(df.copy())
    assert out["MSE"].iloc[0] == pytest.approx((1 + 0 + 1 + 16) / 4)
    assert out["Relative Root MSE"].iloc[0] == pytest.approx(np.sqrt(4.5) / 2)
    assert out["Success Rate"].iloc[0] == pytest.approx(0.75)
    assert out["Estimate Bias"].iloc[0] == pytest.approx(1.0)
    assert out[".50 Cost"].iloc[0] == pytest.approx(25.0)
    only = ev.add_evaluations(df.copy(), only_success=True)
    assert only["MSE"].iloc[0] == pytest.approx((1 + 0 + 16) / 3)
    agg = ev.aggregate_df(out)
    assert len(agg.index) == 1 and agg["Estimate"].iloc[0] == pytest.approx(3.0) and agg["Message"].iloc[0] == "Success"


@pytest.mark.gpu
def test_do_multiple_solves_like_reference_test():
    """test/test_do_multiple_solves.py of the reference, on the engine (same sizes, seeds and arguments)."""
    import rareeventestimation_b200 as ree

    with tempfile.TemporaryDirectory() as d:
        frames = []
        for p in ree.problems_lowdim:
            p.set_sample(1000, seed=5000)
            cbree = ree.CBREE()
            num_runs = 10
            file = ree.do_multiple_solves(p, cbree, num_runs, dir=d, save_other=True, other_list=["t_step"],
                                          addtnl_cols={"col": "val"}, verbose=False)
            df = pd.read_csv(file)
            frames.append(df)
            assert len(df.index) == num_runs, "Dataframe has the wrong number of rows"
            assert list(df.columns) == ["Solver", "Problem", "Seed", "Sample Size", "Truth", "Estimate", "Cost", "Steps",
                                        "Message", "Average Estimate", "Root Weighted Average Estimate",
                                        "VAR Weighted Average Estimate", "CVAR", "SFP", "t_step", "col"][:9] + \
                [c for c in df.columns[9:]]
            assert set(df["Seed"]) == set(range(num_runs)) and (df["col"] == "val").all()
            assert (df["Sample Size"] == 1000).all() and (df["Solver"] == "CBREE").all()
            # the reference's accuracy requirement for these problems (test/test_cbree.py: relative error < 1)
            ok = df[df["Message"] == "Success"]
            assert len(ok.index) >= 8
            assert (abs(ok["Estimate"] - p.prob_fail_true) / p.prob_fail_true < 1.0).all()
        # (load_data cannot read these files in the reference either: without return_other the t_step column holds
        #  "[<NA>]"; the round trip through load_data is covered by the study test below and by the CPU tests)
        ev = ree.add_evaluations(pd.concat(frames, ignore_index=True))
        assert (ev["Relative Root MSE"] < 0.6).all()


@pytest.mark.gpu
def test_study_cbree_observation_window_like_reference_test():
    """test/test_study_cbree_observation_window.py of the reference on the engine (fewer seeds: the replays dominate)."""
    import rareeventestimation_b200 as ree

    with tempfile.TemporaryDirectory() as d:
        for p in ree.problems_lowdim[:2]:
            p.set_sample(5000, seed=5000)
            cbree = ree.CBREE()
            num_runs = 3
            rng_ = range(2, 15)
            file = ree.study_cbree_observation_window(p, cbree, num_runs, dir=d, save_other=True,
                                                      observation_window_range=rng_, other_list=["t_step"],
                                                      addtnl_cols={"col": "val"}, verbose=False)
            df = pd.read_csv(file)
            assert len(df.index) == num_runs * (len(rng_) + 1), "Dataframe has the wrong number of rows"
            full = df[df["observation_window"] == 0]
            assert (~full["divergence_check"].astype(bool)).all() and len(full.index) == num_runs
            replay = df[df["observation_window"] > 0]
            assert replay["divergence_check"].astype(bool).all()
            # a replay never does more steps than the run it replays and costs no new LSF evaluations beyond it
            for s in range(num_runs):
                steps_full = int(full[full["Seed"] == s]["Steps"].iloc[0])
                assert (replay[replay["Seed"] == s]["Steps"] <= steps_full).all()
        # files of this shape go through load_data / add_evaluations / aggregate_df
        cb = ree.CBREE(return_other=True)
        f2 = ree.do_multiple_solves(ree.problems_lowdim[1].set_sample(1000, seed=1), cb, 3, dir=d, prefix="other_",
                                    save_other=True, other_list=["t_step", "sigma", "beta"], verbose=False)
        data = ree.load_data(d, "other_*")
        assert isinstance(data["t_step"][0], np.ndarray) and data["t_step"][0].ndim == 1
        agg = ree.aggregate_df(ree.add_evaluations(data).drop(columns=["index"]))
        assert len(agg.index) == 1 and os.path.exists(f2)


@pytest.mark.gpu
@pytest.mark.needs_reference
def test_reference_harness_unchanged_on_the_engine(tmp_path):
    """SURVEY.md 8f-1 as written: the REFERENCE's own ``do_multiple_solves`` and ``study_cbree_observation_window``
    (evaluation/convergence_analysis.py:34-333, imported unmodified through oracle/ref_loader.py with a plotly stub)
    drive the engine's CBREE on the engine's problems -- what test/test_do_multiple_solves.py:6-24 and
    test/test_study_cbree_observation_window.py:6-28 check for the reference's solver.  Exercises ``set_options(...,
    in_situ=False)`` on solved solvers, ``Solution.other``, ``return_caches`` / ``solve_from_caches`` from foreign code."""
    import pandas as pd

    import ref_loader
    import rareeventestimation_b200 as ree

    ca = ref_loader.load_reference_evaluation()
    num_runs = 10
    for p in ree.problems_lowdim:
        p.set_sample(1000, seed=5000)
        cbree = ree.CBREE(quiet=True)
        file = ca.do_multiple_solves(p, cbree, num_runs, dir=str(tmp_path), save_other=True, other_list=["t_step"],
                                     addtnl_cols={"col": "val"}, verbose=False)
        df = pd.read_csv(file)
        assert len(df.index) == num_runs, "Dataframe has the wrong number of rows"
        assert (df["Message"] == "Success").mean() >= 0.8
        ok = df[df["Message"] == "Success"]
        assert abs(ok["Estimate"].mean() / p.prob_fail_true - 1) < 0.15, p.name
        assert (df["Cost"] == 1000 * (2 + df["Steps"])).all()
    p = ree.problems_lowdim[0].set_sample(1000, seed=5000)
    file = ca.study_cbree_observation_window(p, ree.CBREE(quiet=True), num_runs, dir=str(tmp_path), save_other=True,
                                             observation_window_range=range(2, 15), other_list=["t_step"],
                                             addtnl_cols={"col": "val"}, verbose=False)
    df = pd.read_csv(file)
    assert len(df.index) == num_runs * 14
