"""ctypes binding of libb200aqc.so (C-ABI declared in include/b200aqc.h).

There is deliberately no fallback: if the shared library is missing, or a call returns an
error, an exception is raised.  Every cost evaluation runs on the GPU or not at all.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200aqc.so")

# Every symbol include/b200aqc.h declares (checked by tests/test_abi.py).
SYMBOLS = [
    "b200_abi_version", "b200_last_error", "b200_device_count", "b200_ctx_create",
    "b200_ctx_destroy", "b200_ctx_sync", "b200_ctx_counters", "b200_ctx_last_ms",
    "b200_ctx_set_timing", "b200_ctx_mark", "b200_ctx_elapsed_ms", "b200_ctx_profile",
    "b200_ctx_profile_read", "b200_ctx_profile_sweeps", "b200_sv_alloc", "b200_sv_reserve_slots", "b200_sv_attach", "b200_sv_device_ptr",
    "b200_sv_ipc_export", "b200_sv_ipc_open", "b200_sv_ipc_close", "b200_sv_peer_swap", "b200_sv_peer_swap_strided",
    "b200_sv_num_qubits", "b200_sv_init_zero", "b200_sv_copy", "b200_sv_run",
    "b200_sv_run_inverse", "b200_sv_amp", "b200_sv_expz", "b200_sv_pair_rdm", "b200_sv_pair_rdm_part", "b200_sv_inner", "b200_sv_inner2", "b200_sv_inner2_gather", "b200_sv_gather", "b200_sv_scatter", "b200_sv_gather_ranked",
    "b200_sv_run_inner2", "b200_sv_run_embedded", "b200_sv_run_embedded_inner2", "b200_sv_run_project", "b200_sv_download", "b200_sv_upload", "b200_sv_plan_stats", "b200_sv_plan_detail",
    "b200_mps_create", "b200_mps_destroy", "b200_mps_set_truncation", "b200_mps_num_qubits",
    "b200_mps_init_zero", "b200_mps_set", "b200_mps_bond_dims", "b200_mps_get", "b200_mps_copy",
    "b200_mps_apply", "b200_mps_apply_inverse", "b200_mps_transfer", "b200_mps_amps", "b200_mps_dot",
    "b200_mps_expz", "b200_mps_pair_rdm", "b200_mps_pair_transfer", "b200_mps_stats", "b200_mps_set_chop_rule", "b200_mps_reduce_zeros",
]


class B200Error(RuntimeError):
    """Raised for every non-zero return code of the C-ABI."""


_lib = None


def load():
    """Load libb200aqc.so; raises B200Error if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cu64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64
    dp = ctypes.POINTER(ctypes.c_double)
    L.b200_abi_version.restype = ci
    L.b200_last_error.restype = ctypes.c_char_p
    L.b200_device_count.argtypes = [ctypes.POINTER(ci)]
    L.b200_ctx_create.argtypes = [ci, ctypes.POINTER(vp)]
    L.b200_ctx_destroy.argtypes = [vp]
    L.b200_ctx_sync.argtypes = [vp]
    L.b200_ctx_counters.argtypes = [vp, ctypes.POINTER(cu64)]
    L.b200_ctx_last_ms.argtypes = [vp, dp]
    L.b200_ctx_set_timing.argtypes = [vp, ci]
    L.b200_ctx_mark.argtypes = [vp, ci]
    L.b200_ctx_elapsed_ms.argtypes = [vp, dp]
    L.b200_ctx_profile.argtypes = [vp, ci]
    L.b200_ctx_profile_read.argtypes = [vp, dp, ctypes.POINTER(cu64)]
    L.b200_ctx_profile_sweeps.argtypes = [vp, dp, ci, ctypes.POINTER(ci)]
    L.b200_sv_alloc.argtypes = [vp, ci, ci]
    L.b200_sv_attach.argtypes = [vp, ci, vp]
    L.b200_sv_reserve_slots.argtypes = [vp, ci, ci]
    L.b200_sv_device_ptr.argtypes = [vp, ci, ctypes.POINTER(vp)]
    L.b200_sv_gather.argtypes = [vp, ci, vp, ci, vp]
    L.b200_sv_scatter.argtypes = [vp, ci, vp, ci, vp]
    L.b200_sv_gather_ranked.argtypes = [vp, ci, vp, ci, ci, ci, vp]
    L.b200_sv_ipc_export.argtypes = [vp, ci, ctypes.c_char_p]
    L.b200_sv_ipc_open.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp)]
    L.b200_sv_ipc_close.argtypes = [vp, vp]
    L.b200_sv_peer_swap.argtypes = [vp, ci, ctypes.POINTER(vp), ci, ci]
    L.b200_sv_peer_swap_strided.argtypes = [vp, ci, ctypes.POINTER(vp), ci, ci, vp]
    L.b200_sv_num_qubits.argtypes = [vp, ctypes.POINTER(ci)]
    L.b200_sv_init_zero.argtypes = [vp, ci]
    L.b200_sv_copy.argtypes = [vp, ci, ci]
    L.b200_sv_run.argtypes = [vp, ci, ci, vp, ci, vp, ci]
    L.b200_sv_run_inverse.argtypes = [vp, ci, ci, vp, ci, vp, ci]
    L.b200_sv_amp.argtypes = [vp, ci, cu64, dp]
    L.b200_sv_expz.argtypes = [vp, ci, dp]
    L.b200_sv_pair_rdm.argtypes = [vp, ci, vp, ci, dp]
    L.b200_sv_pair_rdm_part.argtypes = [vp, ci, vp, ci, ci, ci, dp]
    L.b200_sv_inner.argtypes = [vp, ci, ci, ci, dp]
    L.b200_sv_inner2.argtypes = [vp, ci, ci, ci, ci, dp]
    L.b200_sv_run_inner2.argtypes = [vp, ci, ci, vp, ci, vp, ci, ci, ci, ci, ci, dp, ctypes.POINTER(ctypes.c_int)]
    L.b200_sv_run_project.argtypes = [vp, ci, ci, vp, ci, vp, ci, ci, vp, ci, vp, ctypes.POINTER(ctypes.c_int)]
    L.b200_sv_run_embedded.argtypes = [vp, ci, vp, ci, vp, vp, ci, vp, ci, ci]
    L.b200_sv_run_embedded_inner2.argtypes = [vp, ci, vp, ci, vp, vp, ci, vp, ci, ci, ci, ci, ci, dp, ctypes.POINTER(ctypes.c_int)]
    L.b200_sv_inner2_gather.argtypes = [vp, ci, vp, ci, vp, ci, ci, dp]
    L.b200_sv_download.argtypes = [vp, ci, cu64, cu64, vp]
    L.b200_sv_upload.argtypes = [vp, ci, cu64, cu64, vp]
    L.b200_sv_plan_stats.argtypes = [ci, vp, ci, vp, ci, ctypes.POINTER(ctypes.c_int32)]
    L.b200_sv_plan_detail.argtypes = [ci, vp, ci, vp, ci, vp, ci, ctypes.POINTER(ctypes.c_int32)]
    i32p = ctypes.POINTER(ctypes.c_int32)
    L.b200_mps_create.argtypes = [vp, ci, ctypes.c_double, ci, ctypes.POINTER(vp)]
    L.b200_mps_destroy.argtypes = [vp]
    L.b200_mps_set_truncation.argtypes = [vp, ctypes.c_double, ci]
    L.b200_mps_num_qubits.argtypes = [vp, ctypes.POINTER(ci)]
    L.b200_mps_init_zero.argtypes = [vp]
    L.b200_mps_set.argtypes = [vp, vp, vp, vp]
    L.b200_mps_bond_dims.argtypes = [vp, vp]
    L.b200_mps_get.argtypes = [vp, vp, vp]
    L.b200_mps_copy.argtypes = [vp, vp]
    L.b200_mps_apply.argtypes = [vp, vp, ci, vp, ci]
    L.b200_mps_apply_inverse.argtypes = [vp, vp, ci, vp, ci]
    L.b200_mps_transfer.argtypes = [vp, vp, vp, ci, dp]
    L.b200_mps_amps.argtypes = [vp, vp, ci, dp]
    L.b200_mps_dot.argtypes = [vp, vp, dp]
    L.b200_mps_expz.argtypes = [vp, dp]
    L.b200_mps_pair_rdm.argtypes = [vp, vp, ci, dp]
    L.b200_mps_stats.argtypes = [vp, ctypes.POINTER(cu64)]
    L.b200_mps_pair_transfer.argtypes = [vp, vp, vp, ci, dp]
    L.b200_mps_set_chop_rule.argtypes = [ci]
    L.b200_mps_reduce_zeros.argtypes = [dp, ci, ci, ctypes.c_double, ctypes.POINTER(ci), dp]
    del i32p
    for name in SYMBOLS:
        fn = getattr(L, name)
        if name not in ("b200_last_error", "b200_abi_version"):
            fn.restype = ci
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise B200Error(load().b200_last_error().decode("utf-8", "replace"))


def dptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
