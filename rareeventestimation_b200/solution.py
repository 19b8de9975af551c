"""Result container of ``Solver.solve`` -- same attributes as the reference's ``Solution``
(``src/rareeventestimation/solution.py:6-82``)."""
from __future__ import annotations

import numpy as np


class Solution:
    """History of a solve.  ``ensemble_hist`` has shape (S, N, d); for large ensembles it is a
    :class:`~.solver.LazyHistory` that downloads from HBM on access instead of an ndarray."""

    def __init__(self, ensemble_hist, temp_hist, lsf_eval_hist, prob_fail_hist, costs: int, msg: str,
                 num_steps=None, other=None) -> None:
        self.ensemble_hist = ensemble_hist
        self.temp_hist = temp_hist
        self.lsf_eval_hist = lsf_eval_hist
        self.prob_fail_hist = prob_fail_hist
        self.costs = costs
        self.msg = msg
        steps_in_shape, self.N, self.d = ensemble_hist.shape
        self.num_steps = steps_in_shape if num_steps is None else num_steps
        self.other = other

    def get_rel_err(self, prob=None):
        """|estimate - truth| / truth for every entry of ``prob_fail_hist`` (solution.py:69-82)."""
        assert prob is not None, "Need problem to compute rel. error. Please provide the problem!"
        assert prob.prob_fail_true is not None, \
            "Cannot not compute relative error as `Problem.prob_fail_true` is not set."
        self.prob_fail_r_err = abs(np.asarray(self.prob_fail_hist) - prob.prob_fail_true) / prob.prob_fail_true
        return self.prob_fail_r_err
