"""ctypes binding of libfrt2_b200.so (include/frt2.h).  There is no CPU fallback: if the library is
missing or fails to load, importing the product path raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FRT2_LIB") or os.path.join(HERE, "libfrt2_b200.so")   # FRT2_LIB: A/B builds

FRT2_OK = 0
ERR_BAD_ARG, ERR_BAD_DTYPE, ERR_INDEX_OOR, ERR_STATE_OVERFLOW, ERR_CUDA, ERR_MISSING_TENSOR, ERR_NOT_FINALIZED = \
    -1, -2, -3, -4, -5, -6, -7

DBG_TAPS, DBG_GEMM_REF, DBG_ATTN_WARP, DBG_NO_GRAPH, DBG_NO_SKINNY, DBG_NO_LNFOLD = 1, 2, 4, 8, 16, 32
ACT_NONE, ACT_GELU, ACT_POLAR = 0, 1, 2
SLOT_ACTIVE, SLOT_LAST, SLOT_RESET = 1, 2, 4
POOL_MAX_SLOTS = 256
PEER_HANDLE_BYTES = 64
PROF_GEMM, PROF_ATTN_TC, PROF_ATTN_WARP, PROF_LAYER_NORM, PROF_RVQ, PROF_OLA, PROF_GEMM_SKINNY, PROF_ALL = \
    0, 1, 2, 3, 4, 5, 6, -1
PROF_NAMES = {PROF_GEMM: "gemm_tc", PROF_ATTN_TC: "attention_tc", PROF_ATTN_WARP: "attention_warp",
              PROF_LAYER_NORM: "layer_norm", PROF_RVQ: "rvq_gather_sum", PROF_OLA: "overlap_add",
              PROF_GEMM_SKINNY: "gemm_skinny"}


class Frt2Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "rvq_dim", "output_dim", "num_quantizers", "codebook_size", "codebook_dim", "embed_dim",
        "num_layers", "num_heads", "hop_length", "upconv_stride")]


class Frt2EncConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "ssl_in_dim", "ssl_embed_dim", "ssl_out_dim", "ssl_num_layers", "ssl_num_heads", "ssl_ffn_dim", "aco_dim",
        "avg_pooler", "ssl_enc_layers", "ssl_enc_heads", "ssl_enc_ffn_dim", "aco_layers", "aco_heads", "aco_ffn_dim",
        "num_mels", "max_positions")]


class Frt2Error(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"libfrt2_b200 status {status}: {msg}")
        self.status = status


_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); mirrors include/frt2.h one to one
SIGNATURES = {
    "frt2_create": (_i, [C.POINTER(Frt2Config), _i, C.POINTER(_p)]),
    "frt2_load_tensor": (_i, [_p, C.c_char_p, _p, _i, C.POINTER(_i64), _i]),
    "frt2_finalize": (_i, [_p]),
    "frt2_destroy": (None, [_p]),
    "frt2_decode": (_i, [_p, _p, _i, _i64, _i64, _i64, _i, _i, _i, _p, _p, _i64, _p]),
    "frt2_decode_pcm16": (_i, [_p, _p, _i, _i64, _i64, _i64, _i, _i, _i, _p, _p, _i64, _p]),
    "frt2_decode_scatter": (_i, [_p, _p, _i, _i64, _i64, _i64, _i, _i, _i, _p, _p, _i, _p, _p]),
    "frt2_peer_alloc": (_i, [_i, _i64, C.POINTER(_p), C.c_char_p]),
    "frt2_peer_open": (_i, [_i, C.c_char_p, C.POINTER(_p)]),
    "frt2_peer_close": (_i, [_i, _p]),
    "frt2_peer_free": (_i, [_i, _p]),
    "frt2_stream_create": (_i, [_p, _i, _i, C.POINTER(_p)]),
    "frt2_stream_reset": (_i, [_p]),
    "frt2_stream_destroy": (None, [_p]),
    "frt2_stream_tokens": (_i, [_p]),
    "frt2_stream_reserve": (_i, [_p, _i, _i, _i]),
    "frt2_stream_check_error": (_i, [_p, _p, C.POINTER(C.c_int32), _p]),
    "frt2_stream_fetch_errors": (_i, [_p, _p, _p, _p]),
    "frt2_decode_chunk": (_i, [_p, _p, _p, _i, _i64, _i64, _i64, _i, _i, _i, _p, _i64, C.POINTER(_i), _p]),
    "frt2_decode_chunk_pcm16": (_i, [_p, _p, _p, _i, _i64, _i64, _i64, _i, _i, _i, _p, _i64, C.POINTER(_i), _p]),
    "frt2_pool_create": (_i, [_p, _i, _i, C.POINTER(_p)]),
    "frt2_pool_step": (_i, [_p, _p, _p, _i, _i64, _i64, _i, C.POINTER(C.c_int32), _p, _i, _i64,
                            C.POINTER(C.c_int32), _p]),
    "frt2_pool_slot_tokens": (_i, [_p, _i]),
    "frt2_export_state": (_i, [_p, _p, _p, _p, _p, _p, _p, _p]),
    "frt2_import_state": (_i, [_p, _p, _i, _p, _p, _p, _p, _p, _p]),
    "frt2_rvq_encode": (_i, [_p, _p, _i64, _i64, _i64, _i, _i, _i, _i, _p, _p]),
    "frt2_enc_create": (_i, [C.POINTER(Frt2EncConfig), _i, C.POINTER(_p)]),
    "frt2_enc_load_tensor": (_i, [_p, C.c_char_p, _p, _i, C.POINTER(_i64), _i]),
    "frt2_enc_finalize": (_i, [_p]),
    "frt2_enc_destroy": (None, [_p]),
    "frt2_enc_features": (_i, [_p, _p, _p, _i, _i, _p, C.POINTER(_i64), _p]),
    "frt2_enc_audio_features": (_i, [_p, _p, _i64, _i, _i64, _p, _p, _p, _p, C.POINTER(_i64), _p]),
    "frt2_fd_create": (_i, [_p, _i, C.POINTER(_p)]),
    "frt2_fd_load_tensor": (_i, [_p, C.c_char_p, _p, _i, C.POINTER(_i64), _i]),
    "frt2_fd_finalize": (_i, [_p]),
    "frt2_fd_destroy": (None, [_p]),
    "frt2_fd_generate": (_i, [_p, _p, _i, _p, _p, C.c_uint64, _i, _f, _p, _p, _p, C.POINTER(_i64), _p]),
    "frt2_fd_check_error": (_i, [_p, _p]),
    "frt2_resample": (_i, [_i, _p, _i64, _i, _i64, _p, _i, _i, _p, _i64, C.POINTER(_i64), _p]),
    "frt2_decode_resampled": (_i, [_p, _p, _i, _i64, _i64, _i64, _i, _i, _i, _p, _p, _i64, _i, _i, _p, _i64,
                                   C.POINTER(_i64), _p]),
    "frt2_rvq_gather": (_i, [_p, _p, _i, _i64, _i64, _i64, _i, _i, _i, _p, _p, _p]),
    "frt2_set_debug": (_i, [_p, _i]),
    "frt2_get_tap": (_i, [_p, C.c_char_p, _p, _i64, C.POINTER(_i64), _p]),
    "frt2_check_error": (_i, [_p, _p]),
    "frt2_profile": (_i, [_p, _i]),
    "frt2_profile_get": (_i, [_p, _i, C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(C.c_double),
                              C.POINTER(C.c_double)]),
    "frt2_op_gemm": (_i, [_i, _p, _p, _i, _i, _i, _i, _i, _f, _p, _i, _p, _p, _p, _p]),
    "frt2_op_layer_norm": (_i, [_p, _i, _i, _p, _p, _f, _i, _p, _p]),
    "frt2_op_attention": (_i, [_i, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "frt2_op_attention_trace": (_i, [_p]),
    "frt2_op_sample_topk": (_i, [_p, _i, _i, _i, _f, _p, C.c_uint64, _p, _p]),
    "frt2_op_overlap_add": (_i, [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _i, _p]),
    "frt2_last_error": (C.c_char_p, []),
    "frt2_version": (C.c_char_p, []),
}

_lib = None


def load() -> C.CDLL:
    """Load libfrt2_b200.so (built in-tree by fireredtts2_b200/build.py).  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m fireredtts2_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback for the codec decode path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export what frt2.h declares
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().frt2_last_error().decode("utf-8", "replace")


def check(status: int):
    """Map a frt2_status to the exception the reference raises for the same condition."""
    if status == FRT2_OK:
        return
    msg = last_error()
    if status == ERR_INDEX_OOR:
        raise IndexError(msg or "index out of range in self")
    if status == ERR_BAD_DTYPE:
        raise TypeError(msg)
    if status in (ERR_BAD_ARG, ERR_MISSING_TENSOR, ERR_NOT_FINALIZED):
        raise ValueError(msg)
    if status == ERR_STATE_OVERFLOW:
        raise OverflowError(msg)
    raise Frt2Error(status, msg)
