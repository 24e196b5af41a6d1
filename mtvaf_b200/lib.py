"""ctypes binding of the C ABI declared in include/mtvaf_b200.h.

There is NO fallback: if the shared library is missing or a symbol is absent the import raises, and
every call that returns non-zero raises `MtvafError` with the library's message.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libmtvaf_b200.so")
if os.environ.get("MTVAF_LIB_TAG"):          # experiment builds (mtvaf_b200/build.py)
    LIB_PATH = os.path.join(_HERE, "_C", "libmtvaf_b200_%s.so" % os.environ["MTVAF_LIB_TAG"])

F32, BF16 = 0, 1

(EPI_STORE, EPI_GELU, EPI_TANH, EPI_RESID, EPI_ATOMIC_F32, EPI_MUL_DGELU, EPI_MUL_DTANH, EPI_SQNORM, EPI_ROWSCALE,
 EPI_GELU_GRAD, EPI_MUL_AUX) = range(11)


class MtvafError(RuntimeError):
    pass


class Epilogue(C.Structure):
    _fields_ = [("mode", C.c_int32), ("out_dtype", C.c_int32), ("out", C.c_void_p), ("ldo", C.c_int64),
                ("bias", C.c_void_p), ("aux", C.c_void_p), ("ld_aux", C.c_int64), ("out2", C.c_void_p),
                ("ld_out2", C.c_int64), ("rowvec", C.c_void_p), ("alpha", C.c_float), ("p_drop", C.c_float),
                ("seed", C.c_uint64), ("colsum", C.c_void_p)]


_vp, _i, _i64, _f, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64

# name -> argtypes; must list EVERY function declared in include/mtvaf_b200.h (checked by tests)
SIGNATURES = {
    "mtvaf_abi_version": [],
    "mtvaf_device_info": [_vp, _vp, _vp],
    "mtvaf_gemm_bf16": [_vp, _i64, _i, _vp, _i64, _i, _i, _i, _i, C.POINTER(Epilogue), _i, _vp],
    "mtvaf_set_gemm_impl": [_i],
    "mtvaf_set_sm_reserve": [_i],
    "mtvaf_gemm_f32": [_vp, _i64, _i, _vp, _i64, _i, _i, _i, _i, C.POINTER(Epilogue), _i, _vp],
    "mtvaf_skinny_linear_f32": [_vp, _i64, _vp, _i64, _vp, _i, _i, _i, _vp, _i64, _vp],
    "mtvaf_skinny_linear_dgrad": [_vp, _i64, _vp, _i64, _i, _i, _i, _f, _u64, _vp, _i64, _i, _vp, _vp],
    "mtvaf_skinny_linear_wgrad": [_vp, _i64, _vp, _i64, _i, _i, _i, _vp, _i64, _vp],
    "mtvaf_cast_f32_to_bf16": [_vp, _vp, _i64, _vp],
    "mtvaf_cast_bf16_to_f32": [_vp, _vp, _i64, _vp],
    "mtvaf_colsum": [_vp, _i64, _i, _i, _i, _vp, _vp],
    "mtvaf_dropout_apply": [_vp, _vp, _i64, _i, _f, _u64, _vp],
    "mtvaf_rowscale": [_vp, _vp, _vp, _i64, _i, _f, _i, _vp],
    "mtvaf_scale_by_device_scalar": [_vp, _i64, _vp, _vp],
    "mtvaf_add_inplace": [_vp, _i, _vp, _i, _i64, _f, _vp],
    "mtvaf_embed_ln_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp,
                           _vp, _f, _u64, _vp],
    "mtvaf_embed_ln_bwd": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp,
                           _vp, _vp, _f, _u64, _vp],
    "mtvaf_layernorm_fwd": [_vp, _vp, _vp, _vp, _f, _i, _i, _i, _vp, _vp, _vp],
    "mtvaf_layernorm_bwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _u64, _vp],
    "mtvaf_attention_fwd": [_vp, _i64, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _i64, _vp, _vp, _i, _f, _u64, _vp],
    "mtvaf_attention_fwd_ws": [_vp, _i64, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _i64, _vp, _vp, _i, _f, _u64, _vp,
                               _i64, _vp],
    "mtvaf_set_attention_impl": [_i],
    "mtvaf_attention_bwd": [_vp, _i64, _vp, _i64, _vp, _vp, _i, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _vp, _i64, _vp,
                            _vp, _vp, _i, _f, _u64, _vp],
    "mtvaf_attention_bwd_ex": [_vp, _i64, _vp, _i64, _vp, _vp, _i, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _vp, _i64, _vp,
                               _vp, _vp, _i, _f, _u64, _vp, _vp],
    "mtvaf_gate_fwd": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp],
    "mtvaf_gate_bwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp],
    "mtvaf_prompt_grad_combine": [_vp, _vp, _vp, _i, _f, _u64, _i64, _i, _vp, _i, _vp],
    "mtvaf_mean4_fwd": [_vp, _vp, _i64, _i, _i, _i, _vp],
    "mtvaf_mean4_bwd_add": [_vp, _vp, _i64, _i, _i, _vp],
    "mtvaf_softmax_kl_fwd_bwd": [_vp, _i64, _vp, _i, _i, _i, _vp, _vp, _f, _vp],
    "mtvaf_row_sqnorm": [_vp, _i64, _i, _i64, _i, _vp, _vp],
    "mtvaf_probe_labels": [_vp, _vp, _i, _i, _vp],
    "mtvaf_mse_fwd_bwd": [_vp, _vp, _i64, _vp, _vp, _vp],
    "mtvaf_pairwise_sqdist": [_vp, _i64, _i, _i, _i, _i, _vp, _vp],
    "mtvaf_set_pairwise_impl": [_i],
    "mtvaf_pack_features": [_vp, _i64, _vp, _i64, _i, _i, _i, _i64, _vp, _i, _vp],
    "mtvaf_crf_nll_fwd_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp],
    "mtvaf_crf_decode": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "mtvaf_span_offsets": [_vp, _i, _i, _vp, _vp],
    "mtvaf_span_pool_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "mtvaf_span_pool_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "mtvaf_distant_ce_fwd_bwd": [_vp, _i64, _vp, _i, _i, _f, _vp, _vp, _vp],
    "mtvaf_ce_mean_fwd_bwd": [_vp, _vp, _i, _i, _f, _vp, _vp, _vp],
    "mtvaf_combine_loss": [_vp, _i, _vp, _f, _i, _vp, _i, _f, _vp, _vp, _vp],
    "mtvaf_adamw_step": [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _f, _vp, _i, _vp, _vp],
    "mtvaf_adam_dyn_advance": [_vp, _f, _f, _i, _i, _vp],
    "mtvaf_set_step_source": [_vp],
    "mtvaf_advance_step": [_vp, _vp],
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise MtvafError(
            "mtvaf_b200: CUDA extension %s is missing -- run `python -m mtvaf_b200.build` (or "
            "__graft_entry__.build()). There is no CPU / PyTorch fallback for the hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.mtvaf_last_error.restype = C.c_char_p
    lib.mtvaf_last_error.argtypes = []
    lib.mtvaf_launch_count.restype = C.c_uint64
    lib.mtvaf_launch_count.argtypes = []
    lib.mtvaf_attention_fwd_workspace_bytes.restype = C.c_int64
    lib.mtvaf_attention_fwd_workspace_bytes.argtypes = [_i, _i, _i, _i, _i, _i]
    missing = []
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.argtypes = argtypes
        fn.restype = C.c_int
    if missing:
        raise MtvafError("mtvaf_b200: %s does not export %s -- stale build? run `python -m mtvaf_b200.build`"
                         % (LIB_PATH, ", ".join(missing)))
    return lib


_lib = _load()


def last_error() -> str:
    return _lib.mtvaf_last_error().decode("utf-8", "replace")


def call(name: str, *args) -> None:
    rc = getattr(_lib, name)(*args)
    if rc != 0:
        raise MtvafError("%s failed (%d): %s" % (name, rc, last_error()))


def abi_version() -> int:
    return _lib.mtvaf_abi_version()


def raw():
    return _lib
