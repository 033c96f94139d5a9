"""ctypes binding of libfasta_b200.so (the C ABI declared in include/fasta_b200.h).

The library is the ONLY compute path of this package: if it is missing, or no CUDA device is
present, everything that needs arithmetic raises -- there is no CPU fallback.
"""

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libfasta_b200.so")

ABI_VERSION = 1
NSCAL = 64

# scalar-block slots (FB200_S_*)
S_F, S_DX_G0, S_DX_SQ, S_XMXH_SQ, S_PEN, S_RESTART, S_DX_DG, S_DG_SQ, S_G1_SQ, S_THETA = range(10)
S_AUX0, S_AUX1, S_AUX2, S_AUX3 = 10, 11, 12, 13
S_TAU, S_TAU_USED, S_SKIP, S_SKIPPED, S_IT, S_MAXRES, S_G0SQ = 14, 15, 16, 17, 18, 19, 20
S_FRING, FRING = 24, 40

LOSS_NONE, LOSS_LEAST_SQUARES, LOSS_LOGISTIC = 0, 1, 2
PROX_IDENTITY, PROX_SHRINK, PROX_NONNEG, PROX_BOX, PROX_L1BALL, PROX_TV_BALL = range(6)

_p = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_dbl = ctypes.c_double
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/fasta_b200.h one to one
SIGNATURES = {
    "fb200_abi_version": (_int, []),
    "fb200_last_error": (ctypes.c_char_p, []),
    "fb200_workspace_bytes": (_sz, [_i64, _i64]),
    "fb200_dense_uses_tma": (_int, [_p, _i64, _i64, _i64]),
    "fb200_fbs_step": (_int, [_p, _p, _dbl, _int, _dbl, _dbl, _p, _i64, _p, _p, _p, _p, _p, _p]),
    "fb200_forward_step": (_int, [_p, _p, _dbl, _i64, _p, _p]),
    "fb200_step_reduce": (_int, [_p, _p, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "fb200_l1ball_threshold": (_int, [_p, _i64, _dbl, _p, _p, _p]),
    "fb200_accel_step": (_int, [_dbl, _p, _p, _p, _i64, _p, _p, _p, _p, _i64, _int, _int, _p, _p, _p, _p, _p]),
    "fb200_loss_eval": (_int, [_int, _p, _p, _i64, _p, _p, _p, _p]),
    "fb200_bb_reduce": (_int, [_p, _p, _p, _p, _dbl, _i64, _int, _p, _p, _p]),
    "fb200_gemv_loss": (_int, [_p, _i64, _i64, _i64, _p, _int, _p, _p, _p, _p, _p, _sz, _p]),
    "fb200_gemvT_bb": (_int, [_p, _i64, _i64, _i64, _p, _p, _int, _p, _p, _p, _dbl, _p, _p, _sz, _p]),
    "fb200_sweep_supported": (_int, [_p, _i64, _i64, _i64]),
    "fb200_sweep_plan": (_int, [_i64, _i64, ctypes.POINTER(ctypes.c_int)]),
    "fb200_dense_sweep": (_int, [_p, _i64, _i64, _i64, _p, _int, _p, _p, _p, _p, _int, _p, _p, _p, _dbl, _p, _p, _sz, _p]),
    "fb200_dense_sweep_accel": (_int, [_p, _i64, _i64, _i64, _p, _int, _p, _p, _dbl, _p, _p, _p, _p, _int, _p, _p, _p, _dbl,
                                      _p, _p, _sz, _p]),
    "fb200_resident_blocks": (_int, [_i64, _i64]),
    "fb200_resident_scratch_doubles": (_sz, [_i64, _i64]),
    "fb200_resident_cluster_ok": (_int, [_i64, _i64]),
    "fb200_resident_fbs": (_int, [_p, _i64, _i64, _i64, _p, _int, _int, _dbl, _dbl, _dbl] + [_p] * 18 +
                           [_dbl, _dbl, _dbl, _dbl, _int, _int, _int, _int, _int, _int, _int, _int, _int, _p, _p, _p, _p, _p, _p]),
    "fb200_prox_nuclear_scratch_doubles": (_sz, [_i64, _i64]),
    "fb200_prox_nuclear": (_int, [_p, _i64, _i64, _i64, _dbl, _p, _i64, _p, _p, _p, _p]),
    "fb200_gemm_f64": (_int, [_int, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i64, _int, _i64, _p]),
    "fb200_gemm_splits": (_int, [_i64, _i64, _i64]),
    "fb200_ozaki_pad": (_i64, [_i64, _int]),
    "fb200_ozaki_splits": (_int, [_i64, _i64, _i64]),
    "fb200_ozaki_slice_rows": (_int, [_p, _i64, _i64, _i64, _p, _p, _p]),
    "fb200_ozaki_slice_cols": (_int, [_p, _i64, _i64, _p, _i64, _int, _p, _p, _p, _p]),
    "fb200_ozaki_gemm": (_int, [_p, _p, _i64, _p, _p, _i64, _i64, _p, _i64, _p, _int, _i64, _p]),
    "fb200_batched_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "fb200_batched_fbs_step": (_int, [_p, _p, _p, _int, _p, _p, _p, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "fb200_batched_loss": (_int, [_int, _p, _int, _i64, _p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p]),
    "fb200_batched_bb": (_int, [_p, _int, _i64, _p, _p, _p, _p, _p, _int, _i64, _i64, _p, _p, _p, _p]),
    "fb200_batched_select": (_int, [_p, _p, _p, _i64, _i64, _p]),
    "fb200_batched_copy_cols": (_int, [_p, _p, _i64, _i64, _i64, _p, _p, _int, _p]),
    "fb200_tv_grad_nd": (_int, [_p, _p, _int, _p, _p]),
    "fb200_tv_div_nd": (_int, [_p, _p, _int, _p, _p]),
    "fb200_tv_ball_nd": (_int, [_p, _i64, _int, _p, _p]),
    "fb200_tv_div_loss": (_int, [_p, _i64, _i64, _int, _p, _p, _p, _p, _p, _p]),
    "fb200_tv_grad_bb": (_int, [_p, _i64, _i64, _p, _int, _p, _p, _p, _dbl, _p, _p, _p]),
    "fb200_tv_step_div_loss": (_int, [_p, _p, _dbl, _i64, _i64, _int, _p, _p, _p, _p, _p, _p]),
    "fb200_tv_grad_bb_fused": (_int, [_p, _i64, _i64, _p, _int, _p, _p, _p, _dbl, _p, _p, _p]),
    "fb200_snapshot_copy": (_int, [_p, _p, _sz, _p, _p, _p, _p]),
    "fb200_decide_init": (_int, [_p, _dbl, _dbl, _p]),
    "fb200_trial_decide": (_int, [_p, _dbl, _int, _int, _int, _int, _int, _int, _int, _dbl, _int, _dbl, _dbl, _p, _p]),
    "fb200_tv_fista_fused": (_int, [_p, _p, _dbl, _dbl, _i64, _i64, _int] + [_p] * 10),
    "fb200_tv_iter_fused": (_int, [_p, _p, _dbl, _i64, _i64, _int, _p, _p, _p, _p, _p, _p]),
    "fb200_peer_allreduce_bb": (_int, [_p, _int, _i64, _p, _int, _p, _p, _p, _dbl, _int, _p, _p, _p]),
    "fb200_prox_rows": (_int, [_p, _i64, _i64, _int, _dbl, _p, _p, _p]),
    "fb200_dot": (_int, [_p, _p, _i64, _p, _p, _p]),
    "fb200_diff_nrm2sq": (_int, [_p, _p, _i64, _p, _p, _p]),
    "fb200_asum": (_int, [_p, _i64, _p, _p, _p]),
    "fb200_amax": (_int, [_p, _i64, _p, _p]),
    "fb200_prox_apply": (_int, [_p, _int, _dbl, _dbl, _i64, _p, _p, _p]),
    "fb200_sweep_exchange_supported": (_int, []),
    "fb200_dense_sweep_exchange": (_int, [_p, _i64, _i64, _i64, _p, _int, _p, _p, _p, _p, _dbl, _p, _p, _p, _int, _int, ctypes.c_uint32,
                                          _p, _int, _p, _p, _p, _dbl, _p, _p, _p, _p, _p, _sz, _p]),
    "fb200_randn_scratch_bytes": (_sz, [_i64]),
    "fb200_randn_legacy": (_int, [_p, _i64, _p, _p, _sz, _p, _p, _p]),
}

_lib = None


class Fb200Error(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Fb200Error(
            f"{LIB_PATH} not found: build it with `python fasta-python_b200/build.py` "
            "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.fb200_abi_version() != ABI_VERSION:
        raise Fb200Error(f"ABI mismatch: library {lib.fb200_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().fb200_last_error().decode("utf-8", "replace")
        raise Fb200Error(f"{what or 'fb200 call'} failed ({rc}): {msg}")


_torch = None


def require_cuda():
    """torch, once a CUDA device has been seen (the check costs ~5 us per call through torch's NVML probe: remembered)."""
    global _torch
    if _torch is not None:
        return _torch
    import torch
    if not torch.cuda.is_available():
        raise Fb200Error("no CUDA device: fasta-b200 computes only on sm_100a GPUs (no CPU fallback)")
    _torch = torch
    return torch
