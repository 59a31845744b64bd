"""ctypes binding of libipp_b200.so (the C ABI in include/ipp_b200.h).

The library must exist in-tree (built by build.py / __graft_entry__.build()).  There is no fallback:
a missing library or a missing CUDA device is an error at the first GPU call.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IPP_B200_LIB", os.path.join(_HERE, "libipp_b200.so"))  # override: kernel experiments only

_lib = None
_lock = threading.Lock()

_D = C.c_void_p  # device pointer
_I = C.c_int
_LL = C.c_longlong
_F = C.c_double

# name -> argtypes (restype is always int)
SIGNATURES = {
    "ipp_abi_version": [],
    "ipp_gp_layout": [_I, C.POINTER(_LL)],
    "ipp_gram_rbf": [_D, _I, _D, _I, _F, _F, _F, _D, _LL, _D],
    "ipp_gram_kernel": [_I, _D, _I, _D, _I, _F, _F, _F, _D, _LL, _D],
    "ipp_gp_lml_grad": [_D, _D, _I, _D, _D, _D, _D, _D, _I, _I, _D],
    "ipp_gp_fit_final": [_D, _D, _I, _D, _D, _D, _D, _D, _D],
    "ipp_host_chol_plan": [_I, _I, C.POINTER(_I), _LL, C.POINTER(_LL)],
    "ipp_chol_sched_ints": [_I, _I, C.POINTER(_LL)],
    "ipp_gp_lml_grad_batch": [_D, _D, _I, _D, _D, _D, _D, _D, _D, _I, _I, _D, _D, _I, _D, _D],
    "ipp_gp_fit_final_plan": [_D, _D, _I, _D, _D, _D, _D, _D, _D, _D, _I, _D],
    "ipp_gp_posterior_grid": [_D, _D, _I, _F, _F, _F, _F, _F, _F, _I, _F, _F, _I, _LL, _LL, _D, _D, _D, _D, _D, _D, _F, _I, _D],
    "ipp_gp_posterior_points": [_D, _D, _I, _F, _F, _F, _F, _D, _LL, _D, _D, _D, _D, _I, _D],
    "ipp_gp_posterior_points_tabled": [_D, _D, _I, _F, _F, _F, _F, _D, _LL, _D, _D, _D, _D, _D, _LL, _D, _F, _I, _D],
    "ipp_mi_logdet": [_D, _I, _D, _I, _F, _F, _F, _D, _D, _D, _D, _D],
    "ipp_count_close": [_D, _I, _D, _I, _F, _D, _D, _D, _D, _D],
    "ipp_node_terms": [_I, _D, _I, _D, _I, _I, _D, _D, _I, _D, _D, _D],
    "ipp_branch_gain": [_I, _D, _I, _D, _D, _D, _D, _D, _D, _D, _I, _I, _D, _I, _D, _D, _I, _D, _D, _D],
    "ipp_leaf_penalty": [_D, _D, _D, _I, _D, _I, _D, _F, _D, _D],
    "ipp_ucb_visit_penalty": [_D, _I, _D, _I, _F, _D, _D],
    "ipp_argmax_first": [_D, _I, _D, _D, _D],
    "ipp_ground_truth_grid": [_D, _I, _D, _I, _F, _F, _I, _F, _F, _I, _F, _F, _D, _D],
    "ipp_grid_metrics": [_D, _D, _D, _LL, _D, _D, _D],
    "ipp_poisson_loglik": [_D, _I, _I, _D, _D, _D, _I, _F, _D, _D],
    "ipp_host_norm2": [_I, _D, _I, _D],
    "ipp_host_leaves_dfs": [_D, _I, _D, _D],
    "ipp_host_tree_grow": [_I, _I, _D, _D, _D, _D, _D, _I, _I, _D, _I, _D, _F, _F, _F, _D, _D, _I, _D, _D, _D, _D, _D],
}


class IppError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach signatures for every symbol the header declares."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise IppError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a).  This package has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = _I
        _lib = lib
        return lib


def check(rc: int, what: str):
    if rc != 0:
        raise IppError(f"{what} failed with status {rc} ({'bad argument' if rc == -1 else 'CUDA error'})")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise IppError("no CUDA device visible: the B200 path has no CPU fallback")
    return torch


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
