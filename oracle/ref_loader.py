"""TEST INFRASTRUCTURE ONLY -- loader for the *live* reference package.

Imports the unmodified reference modules from ``/root/reference/src`` (build
container) or from ``oracle/_ref/src`` (the git-ignored copy ``oracle/make_ref.sh``
makes; it travels to the GPU box with gpurun) without executing the
reference's ``__init__`` (which pulls plotly / skfem).
Used by ``tests/golden/make_golden.py`` to generate the committed golden
fixtures, by the ``needs_reference`` tests that pin ``oracle/`` against the
reference itself, and by ``bench.py --impl reference`` / its ``cpu_baseline`` leg
to time the reference's own CBREE.  Nothing in ``rareeventestimation_b200`` may
import this.
"""
import importlib
import os
import sys
import types

_PKG = "rareeventestimation"


def _find_reference_src() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("REE_REFERENCE_SRC"), "/root/reference/src", os.path.join(here, "_ref", "src")):
        if cand and os.path.isdir(os.path.join(cand, _PKG)):
            return cand
    return "/root/reference/src"


REF_SRC = _find_reference_src()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, _PKG))


def _stub(name, **attrs):
    if name not in sys.modules:
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
    return sys.modules[name]


def load_reference():
    """Return a namespace with the reference's solver/problem/utilities modules."""
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REF_SRC}")
    import numpy as np

    if not hasattr(np, "int"):
        np.int = int  # diffusion.py:133 uses the alias removed in NumPy 1.24
    try:  # only needed for verbose tables (solver.py:401-406)
        import prettytable  # noqa: F401
    except ImportError:

        class PrettyTable:  # minimal stand-in, never used with verbose=False
            def __init__(self, *a, **k):
                self.rows = []

            def add_row(self, r):
                self.rows.append(r)

            def get_string(self, **k):
                return str(self.rows[-1]) if self.rows else ""

        _stub("prettytable", PrettyTable=PrettyTable)

    root = os.path.join(REF_SRC, _PKG)
    if _PKG not in sys.modules or not getattr(sys.modules[_PKG], "_ree_stub", False):
        parent = types.ModuleType(_PKG)
        parent.__path__ = [root]
        parent._ree_stub = True
        sys.modules[_PKG] = parent
    ns = types.SimpleNamespace()
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ns.solver = importlib.import_module(_PKG + ".solver")
        ns.utilities = importlib.import_module(_PKG + ".utilities")
        ns.problem = importlib.import_module(_PKG + ".problem.problem")
        ns.toy = importlib.import_module(_PKG + ".problem.toy_problems")
        ns.solution = importlib.import_module(_PKG + ".solution")
        ns.sis_acs = importlib.import_module(_PKG + ".sis.SIS_aCS")
    return ns


def load_reference_diffusion():
    """The PDE module is slow to import (KL root finding); load it on demand."""
    load_reference()
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module(_PKG + ".problem.diffusion")


def load_reference_evaluation():
    """The reference's ``evaluation/convergence_analysis.py`` (do_multiple_solves, load_data, add_evaluations,
    aggregate_df).  Its package pulls plotly through ``evaluation/constants.py``; a stub with the three names that
    module touches at import time is enough."""
    load_reference()

    class _Templates(dict):
        def __missing__(self, key):
            self[key] = _Templates()
            return self[key]

    if "plotly" not in sys.modules:
        colors = types.SimpleNamespace(qualitative=types.SimpleNamespace(D3=["#1f77b4"] * 10, Plotly=["#636efa"] * 10))
        _stub("plotly", io=None, express=None)
        pio = _stub("plotly.io", templates=_Templates())
        px = _stub("plotly.express", colors=colors)
        sys.modules["plotly"].io, sys.modules["plotly"].express = pio, px
    return importlib.import_module(_PKG + ".evaluation.convergence_analysis")
