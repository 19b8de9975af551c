"""rareeventestimation_b200 -- B200-native engine for the CBREE hot path of ``rareeventestimation``.

Drop-in for the path ``ree.CBREE().solve(problem)`` with ``NormalProblem`` / ``Solution`` /
``prob_fail_hist`` / caches (reference: AlthausKonstantin/RareEventEstimation v0.3.0).  All N-sized work runs
in hand-written sm_100a CUDA kernels behind ``libree_b200.so`` (C ABI: ``include/ree_b200.h``).
No CPU fallback: importing is cheap and GPU-free, but every solve needs the built extension and a B200.
"""
from ._lib import ReeError  # noqa: F401
from .problem import DeviceSample, NormalProblem, Problem, Vectorizer  # noqa: F401
from .lsf import AffineLSF, ConvexLSF, DeviceLSF, DiffusionLSF, FujitaRackwitzLSF, OscillatorLSF  # noqa: F401
from .solution import Solution  # noqa: F401
from .mixture import MixtureModel, VMFNMixture  # noqa: F401
from .solver import CBREE, CBREECache, Solver, flatten_cache_list, get_slope  # noqa: F401
from .baselines import CMC, SIS  # noqa: F401
from .utilities import gaussian_logpdf, importance_sampling, my_log_cvar, my_softmax  # noqa: F401
from .evaluation import (  # noqa: F401
    add_evaluations, aggregate_df, do_multiple_solves, load_data, study_cbree_observation_window,
)
from .toy_problems import (  # noqa: F401
    lsf_convex, lsf_lin, lsf_nonlin_osc, make_fujita_rackwitz, make_linear_problem, prob_convex, prob_nonlin_osc,
    problems_highdim, problems_lowdim,
)

__version__ = "0.1.0"


def __getattr__(name):
    # the diffusion problem builds its KL table at first use (2-3 s of scalar root finding)
    if name in ("diffusion_problem", "diffusion"):
        import importlib

        _d = importlib.import_module(__name__ + ".diffusion")

        return _d.diffusion_problem if name == "diffusion_problem" else _d
    raise AttributeError(name)
