"""ctypes binding of ``libree_b200.so`` (the C ABI declared in ``include/ree_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (``make -C rareeventestimation_b200/csrc``).
There is no CPU fallback: if the shared object is missing or a call fails, a :class:`ReeError` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libree_b200.so")

TGT_KINDS = {"algebraic": 0, "tanh": 1, "arctan": 2, "erf": 3, "relu": 4, "sigmoid": 5}
LSF_EXTERNAL, LSF_AFFINE, LSF_CONVEX, LSF_FUJITA_RACKWITZ, LSF_OSCILLATOR, LSF_DIFFUSION = range(6)
NOISE_INJECT, NOISE_PHILOX = 0, 1
ABI_VERSION = 9


class ReeError(RuntimeError):
    """A call into libree_b200.so failed (CUDA error or bad argument)."""


class ReeLsf(C.Structure):
    """struct ree_lsf (include/ree_b200.h)."""

    _fields_ = [
        ("kind", C.c_int32),
        ("d", C.c_int32),
        ("coef", C.c_void_p),
        ("offset", C.c_double),
        ("p0", C.c_double),
        ("p1", C.c_double),
        ("n_ele", C.c_int32),
        ("reserved", C.c_int32),
    ]


class ReeAdapt(C.Structure):
    """struct ree_adapt (include/ree_b200.h): the sigma / beta searches of one iteration."""

    _fields_ = [
        ("tgt_kind", C.c_int32),
        ("do_sigma", C.c_int32),
        ("do_beta", C.c_int32),
        ("maxiter", C.c_int32),
        ("maxfev", C.c_int32),
        ("use_peers", C.c_int32),
        ("sigma", C.c_double),
        ("sigma_hi", C.c_double),
        ("cvar_tgt", C.c_double),
        ("xatol", C.c_double),
        ("beta", C.c_double),
        ("ess_tgt_n", C.c_double),
        ("xtol", C.c_double),
        ("factor", C.c_double),
        ("epsfcn", C.c_double),
    ]


SCALAR_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)

_P = C.c_void_p
_I64 = C.c_int64
_U64 = C.c_uint64
_INT = C.c_int
_DBL = C.c_double
_LSFP = C.POINTER(ReeLsf)

# name -> (restype, argtypes); every symbol declared in include/ree_b200.h
PROTOTYPES = {
    "ree_abi_version": (_INT, []),
    "ree_last_error": (C.c_char_p, []),
    "ree_num_ctas": (_INT, []),
    "ree_launch_count": (_I64, []),
    "ree_fused_path": (_INT, [_INT]),
    "ree_max_dim": (_INT, []),
    "ree_workspace_bytes": (_I64, [_INT]),
    "ree_fetch_small": (_INT, [_P, _INT, _P, _P]),
    "ree_download_vector": (_INT, [_P, _P, _I64, _P]),
    "ree_vmfn_fit_sums": (_INT, [_P, _I64, _INT, _I64, _P, _P, _P, _P, _P]),
    "ree_vmfn_sample": (_INT, [_P, _I64, _INT, _I64, _P, _DBL, _DBL, _DBL, _DBL, _U64, _U64, _I64, _P, _P, _P, _P]),
    "ree_vmfn_logpdf": (_INT, [_P, _I64, _INT, _I64, _P, _DBL, _DBL, _DBL, _DBL, _P, _P, _P, _P, _P, _LSFP, _P]),
    "ree_cbree_adapt": (_INT, [_P, _P, _I64, C.POINTER(ReeAdapt), _P, _P, _P, _P]),
    "ree_scalar_brent": (_INT, [SCALAR_FN, _P, _DBL, _DBL, _DBL, _INT, _P, _P, _P, _P]),
    "ree_scalar_hybr1": (_INT, [SCALAR_FN, _P, _DBL, _DBL, _INT, _DBL, _DBL, _P, _P, _P, _P]),
    "ree_aos_to_soa": (_INT, [_P, _P, _I64, _INT, _I64, _P]),
    "ree_soa_to_aos": (_INT, [_P, _P, _I64, _INT, _I64, _P]),
    "ree_upload_ensemble": (_INT, [_P, _P, _I64, _INT, _I64, _P, _I64, _P]),
    "ree_download_ensemble": (_INT, [_P, _P, _I64, _INT, _I64, _P, _I64, _P]),
    "ree_philox_normal": (_INT, [_P, _I64, _INT, _I64, _U64, _U64, _I64, _P]),
    "ree_lsf_eval": (_INT, [_LSFP, _P, _I64, _I64, _P, _P]),
    "ree_energy": (_INT, [_P, _I64, _INT, _I64, _P, _P]),
    "ree_logtgt": (_INT, [_P, _P, _I64, _DBL, _INT, _P, _P]),
    "ree_reduce_ess": (_INT, [_P, _P, _I64, _DBL, _DBL, _INT, _P, _P, _P, _P]),
    "ree_reduce_cvar": (_INT, [_P, _I64, _DBL, _DBL, _INT, _P, _P, _P]),
    "ree_reduce_scaled": (_INT, [_P, _I64, _DBL, _P, _P, _P, _P]),
    "ree_vector_range": (_INT, [_P, _I64, _P, _P, _P]),
    "ree_peer_create": (_INT, [_INT, _INT, _P]),
    "ree_peer_connect": (_INT, [_P]),
    "ree_peer_connect_local": (_INT, [C.POINTER(C.c_int)]),
    "ree_peer_allreduce_sum": (_INT, [_P, _I64, _P]),
    "ree_peer_allreduce_max": (_I64, []),
    "ree_peer_ready": (_INT, []),
    "ree_peer_combine_expsums": (_INT, [_P, _P]),
    "ree_peer_check": (_INT, []),
    "ree_merge_expsums": (_INT, [_P, _INT, _P, _P]),
    "ree_count_failed": (_INT, [_P, _I64, _P, _P, _P]),
    "ree_cbree_step": (_INT, [_P, _P, _I64, _INT, _I64, _DBL, _P, _P, _INT, _P, _U64, _U64, _I64, _LSFP, _P, _P, _P,
                              _P, _P, _P]),
    "ree_sums_len": (_I64, [_INT]),
    "ree_cbree_moments": (_INT, [_P, _I64, _INT, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ree_ensemble_sums": (_INT, [_P, _I64, _INT, _I64, _P, _P, _P, _P, _P]),
    "ree_moments_chol": (_INT, [_P, _P, _INT, _INT, _INT, _P, _P, _P, _P, _P, _P]),
    "ree_chol_scaled": (_INT, [_P, _INT, _DBL, _P, _P, _P]),
    "ree_cbree_maha": (_INT, [_P, _I64, _INT, _I64, _P, _P, _P, _P, _P, _DBL, _DBL, _INT, _P, _P, _P, _P, _P, _P,
                              _P]),
    "ree_sis_reduce": (_INT, [_P, _I64, _DBL, _DBL, _INT, _P, _P, _P, _P]),
    "ree_scan_scratch_doubles": (_I64, [_I64]),
    "ree_scan_inclusive": (_INT, [_P, _I64, _P, _P, _P]),
    "ree_sis_resample": (_INT, [_P, _I64, _I64, _U64, _U64, _I64, _P, _P]),
    "ree_gather_columns": (_INT, [_P, _INT, _I64, _P, _I64, _P, _I64, _P, _P, _P]),
    "ree_acs_propose": (_INT, [_P, _I64, _INT, _I64, _DBL, _DBL, _U64, _U64, _I64, _P, _P]),
    "ree_acs_accept": (_INT, [_P, _P, _P, _P, _I64, _INT, _I64, _DBL, _U64, _U64, _I64, _P, _I64, _I64, _P, _P, _P,
                              _P]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes.  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ReeError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C rareeventestimation_b200/csrc). "
            "This package has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.ree_abi_version() != ABI_VERSION:
        raise ReeError(f"ABI mismatch: library {lib.ree_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().ree_last_error().decode("utf-8", "replace")
        raise ReeError(f"{what or 'libree_b200'} failed (code {rc}): {msg}")
