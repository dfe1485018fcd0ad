"""ctypes binding of libgdmcf_sm100.so (C ABI declared in include/gdmcf_sm100.h).

PyTorch is only the allocator and stream owner here: every call passes `tensor.data_ptr()` and the
current CUDA stream handle. There is no fallback: if the library is missing or a call fails, the
caller gets an exception (GdmcfError / AssertionError-compatible), never a silent torch path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgdmcf_sm100.so")

MAX_SEG = 8
EPI_STORE, EPI_BIAS_ACT, EPI_COSINE = 0, 1, 2
ACT_NONE, ACT_TANH, ACT_RELU = 0, 1, 2


class GdmcfError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p * MAX_SEG),
        ("b", C.c_void_p * MAX_SEG),
        ("lda", C.c_int64 * MAX_SEG),
        ("ldb", C.c_int64 * MAX_SEG),
        ("k", C.c_int32 * MAX_SEG),
        ("n_seg", C.c_int32),
        ("m", C.c_int32),
        ("n", C.c_int32),
    ]


class Epilogue(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("act", C.c_int32),
        ("alpha", C.c_float),
        ("t_const", C.c_int32),
        ("out_f32", C.c_void_p),
        ("out_bf16", C.c_void_p),
        ("out_bf16_lo", C.c_void_p),
        ("ld_f32", C.c_int64),
        ("ld_bf16", C.c_int64),
        ("bias", C.c_void_p),
        ("ld_bias", C.c_int64),
        ("row_t", C.c_void_p),
        ("row_scale", C.c_void_p),
        ("col_scale", C.c_void_p),
        ("c1", C.c_void_p),
        ("c2", C.c_void_p),
        ("xt", C.c_void_p),
        ("ld_xt", C.c_int64),
    ]


class Refresh(C.Structure):
    _fields_ = [
        ("cols_used", C.c_int32),
        ("n_tcols", C.c_int32),
        ("hi", C.c_void_p), ("lo", C.c_void_p), ("ld_hi", C.c_int64),
        ("t_hi", C.c_void_p), ("t_lo", C.c_void_p), ("ld_t", C.c_int64),
        ("inv_norm", C.c_void_p),
        ("delta", C.c_void_p), ("ld_delta", C.c_int64), ("base", C.c_void_p),
        ("tcols", C.c_void_p),
        ("rowpart", C.c_void_p),
        ("row_coef", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/gdmcf_sm100.h declares.
_P, _I, _L, _F, _U64, _SZ = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64, C.c_size_t
SIGNATURES = {
    "gdmcf_last_error": (C.c_char_p, []),
    "gdmcf_abi_version": (_I, []),
    "gdmcf_device_check": (_I, []),
    "gdmcf_num_sms": (_I, []),
    "gdmcf_launch_count": (C.c_ulonglong, []),
    "gdmcf_spmm_plan": (_I, [_P, _I, _I, _P, _I, _P, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "gdmcf_spmm_csr_f32": (_I, [_P, _P, _P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _F, _F, _P]),
    "gdmcf_lightgcn_propagate_f32": (_I, [_P, _P, _P, _I, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "gdmcf_build_norm_adj": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "gdmcf_lightgcn_propagate_sym_f32": (_I, [_P, _P, _P, _I, _P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "gdmcf_norm_adj_dinv": (_I, [_P, _P, _I, _I, _P, _P]),
    "gdmcf_gemm_auto_splits": (_I, [_I, _I, _I]),
    "gdmcf_gemm_set_sm_limit": (_I, [_I]),
    "gdmcf_gemm_workspace_bytes": (_SZ, [_I, _I, _I]),
    "gdmcf_gemm_bf16_tn": (_I, [C.POINTER(GemmDesc), C.POINTER(Epilogue), _I, _P, _SZ, _P]),
    "gdmcf_cast_bf16": (_I, [_P, _L, _P, _P, _L, _I, _I, _P]),
    "gdmcf_cast_bf16_transpose": (_I, [_P, _L, _P, _P, _L, _I, _I, _P]),
    "gdmcf_densify_rows": (_I, [_P, _P, _P, _I, _I, _P, _L, _P, _L, _P]),
    "gdmcf_qsample_dropout": (_I, [_P, _L, _P, _I, _P, _P, _P, _P, _F, _U64, _U64, _P, _P, _L, _P, _P, _L, _I, _I, _P]),
    "gdmcf_onehot_noise": (_I, [_P, _L, _P, _F, _F, _P, _P, _U64, _U64, _P, _P, _L, _I, _I, _P]),
    "gdmcf_lightgcn_hot_rows": (_I, []),
    "gdmcf_lightgcn_propagate_bf16": (_I, [_P, _P, _P, _I, _I, _P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "gdmcf_scale_cols_cast": (_I, [_P, _L, _P, _P, _P, _L, _I, _I, _P]),
    "gdmcf_graph_noise_step": (_I, [_P, _L, _P, _I, _I, _F, _I, _U64, _U64, _P, _P, _P, _I, _I, _P]),
    "gdmcf_encode_onehot_gather_workspace_bytes": (_SZ, [_I, _I]),
    "gdmcf_encode_onehot_gather": (_I, [_P, _P, _P, _I, _P, _P, _L, _I, _P, _L, _P, _SZ, _P, _P]),
    "gdmcf_onehot_tables": (_I, [_P, _L, _I, _I, _P, _P, _L, _P]),
    "gdmcf_time_bias_table": (_I, [_P, _P, _P, _L, _P, _I, _I, _I, _P, _P, _L, _P]),
    "gdmcf_bias_act_rows": (_I, [_P, _L, _P, _L, _P, _I, _I, _P, _L, _P, _P, _L, _I, _I, _P]),
    "gdmcf_gather_rows": (_I, [_P, _L, _P, _P, _L, _P, _P, _L, _I, _I, _P]),
    "gdmcf_sgemm_small": (_I, [_P, _L, _I, _P, _L, _I, _P, _L, _I, _I, _I, _F, _F, _P]),
    "gdmcf_colsum_f32": (_I, [_P, _L, _I, _I, _P, _P]),
    "gdmcf_mix_rownorm": (_I, [_P, _L, _P, _L, _P, _P, _L, _P, _P, _L, _P, _I, _I, _P]),
    "gdmcf_row_inv_norm": (_I, [_P, _L, _P, _I, _I, _P]),
    "gdmcf_mask_topk": (_I, [_P, _L, _I, _I, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "gdmcf_topn_metrics": (_I, [_P, _I, _I, _P, _P, _P, _P, _I, _P, _P]),
    "gdmcf_colsum_f64": (_I, [_P, _I, _I, _P, _P]),
    "gdmcf_mse_rows": (_I, [_P, _L, _P, _L, _I, _I, _P, _P]),
    "gdmcf_adamw_fused": (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _P, _F, _P]),
    "gdmcf_counter_add": (_I, [_P, _U64, _P]),
    "gdmcf_user_tower_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "gdmcf_user_tower": (_I, [_P, _L, _P, _L, _P, _L, _P, _P, _L, _P, _P, _I, _I, _I, _I, _P, _L, _P, _P, _P, _L, _P, _L, _P, _SZ,
                              _P, _I, _P]),
    "gdmcf_adamw_rows_lazy": (_I, [_P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _F, _F, _F, _F, _F, _I, _P, _F, _P]),
    "gdmcf_adamw_partitioned": (_I, [_P, _P, _L, _P, _P, _I, _I, _F, _F, _F, _F, _F, _I, _P, _F, _P, _I, _P]),
    "gdmcf_adamw_refresh_splits": (_I, [_I, _I]),
    "gdmcf_adamw_refresh": (_I, [_P, _P, _L, _P, _P, _I, _I, _F, _F, _F, _F, _F, _I, _P, _F, C.POINTER(Refresh), _P]),
    "gdmcf_loss_grad": (_I, [_P, _L, _P, _L, _P, _P, _P, _I, _P, _P, _L, _P, _P, _L, _P, _P, _I, _I, _P]),
    "gdmcf_transpose_bf16": (_I, [_P, _L, _P, _L, _I, _I, _P]),
    "gdmcf_ew_binary": (_I, [_I, _P, _L, _P, _L, _F, _F, _P, _L, _P, _P, _L, _I, _I, _P]),
    "gdmcf_mix_backward": (_I, [_P, _L, _P, _L, _P, _L, _P, _P, _L, _P, _L, _P, _I, _I, _P]),
    "gdmcf_ntxent_rows": (_I, [_P, _L, _I, _F, _F, _P, _P, _P, _L, _P]),
    "gdmcf_scatter_rows_add": (_I, [_P, _L, _P, _P, _L, _I, _I, _P]),
    "gdmcf_lt_history_update": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "gdmcf_loss_terms": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "gdmcf_sample_timesteps": (_I, [_P, _P, _I, _I, _I, C.c_double, _U64, _U64, _P, _P, _P, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (no compute). Raises GdmcfError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GdmcfError(
            f"{LIB_PATH} is missing: build it with `python -m gdmcf_b200.build` "
            "(the engine has no non-CUDA code path)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI drifted from the header
        fn.restype = res
        fn.argtypes = args
    if lib.gdmcf_abi_version() != 1:
        raise GdmcfError("libgdmcf_sm100.so ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load().gdmcf_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = f"{what} failed (rc={rc}): {last_error()}"
    if rc == -1:
        raise AssertionError(msg)  # the reference asserts on shape errors
    raise GdmcfError(msg)


def ptr(t) -> int | None:
    if t is None:
        return None
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise GdmcfError("gdmcf_b200 kernels need CUDA tensors (no CPU fallback)")
