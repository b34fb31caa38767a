"""ctypes binding of liblgcn_b200.so (the C ABI in include/lgcn_b200.h).

There is NO fallback: if the library is missing or an entry point fails, a
`LgcnLibraryError` is raised.  The product never routes around the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("LGCN_B200_LIB", PKG / "liblgcn_b200.so"))  # override: tuning variants

ABI_VERSION = 4
F32, BF16, F16 = 0, 1, 2
HUB_DEG = 256
SEG_EDGES = 1024
MAX_PEERS = 8
ERR_INVALID_ARG = -1
ERR_UNSUPPORTED = -2


class LgcnLibraryError(RuntimeError):
    pass


class GraphStruct(C.Structure):
    """lgcn_graph_t"""
    _fields_ = [
        ("n_nodes", C.c_int64), ("nnz", C.c_int64),
        ("rowptr", C.c_void_p), ("col", C.c_void_p), ("dinv", C.c_void_p),
        ("light_desc", C.c_void_p), ("n_light", C.c_int64),
        ("seg_row", C.c_void_p), ("seg_begin", C.c_void_p), ("seg_len", C.c_void_p),
        ("seg_hub", C.c_void_p), ("n_seg", C.c_int64),
        ("hub_seg0", C.c_void_p), ("hub_nseg", C.c_void_p), ("hub_counter", C.c_void_p),
        ("n_hub", C.c_int64), ("partial", C.c_void_p),
    ]


class LayerArgs(C.Structure):
    """lgcn_layer_args_t"""
    _fields_ = [
        ("d", C.c_int), ("src_dtype", C.c_int), ("dst_dtype", C.c_int), ("scale_src", C.c_int),
        ("src", C.c_void_p), ("dst", C.c_void_p), ("base", C.c_void_p),
        ("acc_in", C.c_void_p), ("acc_out", C.c_void_p), ("acc_scale", C.c_float),
        ("grad_mode", C.c_int), ("inv_layers", C.c_float), ("reg_coef", C.c_float),
        ("cnt", C.c_void_p), ("emb", C.c_void_p), ("grad", C.c_void_p),
        ("adam_m", C.c_void_p), ("adam_v", C.c_void_p), ("adam_hp", C.c_void_p),
        ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
        ("zero_base", C.c_int),
        ("n_dst_peers", C.c_int), ("dst_row_offset", C.c_int64), ("dst_peers", C.c_void_p * MAX_PEERS),
        ("src_scale", C.c_void_p), ("dst_scale", C.c_void_p), ("edge_w", C.c_void_p),
        ("push_emb", C.c_int), ("dst_multicast", C.c_void_p), ("dst_route_rows", C.c_int64),
    ]


# name -> (restype, argtypes); every symbol include/lgcn_b200.h declares
_P, _I64, _I, _F, _D = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double
SIGNATURES = {
    "lgcn_abi_version": (C.c_int, []),
    "lgcn_last_error": (C.c_char_p, []),
    "lgcn_propagate_layer": (C.c_int, [C.POINTER(GraphStruct), C.POINTER(LayerArgs), _P]),
    "lgcn_reduce_rows": (C.c_int, [_P, _I, _I64, _I64, _P, C.POINTER(LayerArgs), _P]),
    "lgcn_scale_rows_push": (C.c_int, [_P, _P, _I64, _I, _I, C.POINTER(C.c_void_p), _I, _I64, _P]),
    "lgcn_exchange_rows_push": (C.c_int, [_P, _P, _I, _P, _I64, _I64, _I, C.POINTER(C.c_void_p), _I, _P]),
    "lgcn_bpr_fwd_bwd": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I, _F, _F, _P, _P, _P, _P, _P, _P]),
    "lgcn_padded_ids": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _P, _P, _I, _I64, _P, _P, _P]),
    "lgcn_bpr_fwd_bwd_rows": (C.c_int, [_P, _P, _I64, _I, _I64, _I, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
    "lgcn_zero_rows": (C.c_int, [_P, _I, _P, _I64, _P]),
    "lgcn_ssm_fwd_bwd": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I, _I64, _I64, _I, _F, _F, _F, _P, _P, _P, _P, _P, _P]),
    "lgcn_adam_tick": (C.c_int, [_P, _P, _D, _D, _D, _P]),
    "lgcn_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _P, _D, _D, _D, _P]),
    "lgcn_uniform_sample": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _I64, _I, C.c_uint64, C.c_uint32, _P, _P, _P]),
    "lgcn_uniform_sample_weighted": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I, C.c_uint64, C.c_uint32,
                                               _P, _P, _P]),
    "lgcn_compact_triples": (C.c_int, [_P, _P, _I64, _P, _P, _P, _P]),
    "lgcn_score_topk": (C.c_int, [_P, _P, _P, _I64, _I64, _I, _P, _P, _I, _F, _I, _P, _P, _P, _I64, _P]),
    "lgcn_score_topk_workspace_bytes": (C.c_int64, [_I64, _I64, _I, _I]),
    "lgcn_score_topk_debug": (C.c_int, [_P, _P, _P, _I64, _I64, _I, _P, _P, _I, _F, _I, _P, _P, _P, _I64, _P, _P]),
    "lgcn_score_dense_f32": (C.c_int, [_P, _P, _P, _I64, _I64, _I, _P, _P]),
    "lgcn_ingest_tiles": (C.c_int64, [_I64]),
    "lgcn_ingest_count": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P]),
    "lgcn_ingest_emit": (C.c_int, [_P, _I64, _P, _P, _P, _I64, _P, _P, _P, _I64, _P, _P]),
    "lgcn_rank_metrics": (C.c_int, [_P, _I64, _I, _P, _P, _P, C.POINTER(C.c_int32), _I, _P, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library and bind every declared symbol (loudly)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise LgcnLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m furusato_recommend_b200.build` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise LgcnLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.lgcn_abi_version() != ABI_VERSION:
        raise LgcnLibraryError(
            f"ABI mismatch: library {lib.lgcn_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().lgcn_last_error()
        raise LgcnLibraryError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
