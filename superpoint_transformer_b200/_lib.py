"""ctypes binding of libspt_b200.so (C ABI declared in include/spt_b200.h).

There is NO CPU or eager-PyTorch fallback: if the CUDA library cannot be loaded
the import of any op fails loudly (RuntimeError), and every op refuses non-CUDA
tensors.  The library is built in-tree by `csrc/build.py` (nvcc, sm_100a).
"""
import ctypes
import os
import threading

from .csrc import build as _build

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_f32 = ctypes.c_float
c_ptr = ctypes.c_void_p
c_size = ctypes.c_size_t

_LOCK = threading.Lock()
_LIB = None

# name -> (restype, argtypes); mirrors include/spt_b200.h one to one
SIGNATURES = {
    "spt_abi_version": (c_int, []),
    "spt_last_error": (ctypes.c_char_p, []),
    "spt_build_info": (ctypes.c_char_p, []),
    "spt_group_index_workspace_bytes": (c_size, [c_i64, c_i64]),
    "spt_group_index": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr,
                                c_ptr, c_size, c_ptr]),
    "spt_invert_permutation": (c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    "spt_gather_i32": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "spt_expand_pointers_i32": (c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    "spt_segment_sum_i64": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "spt_gather_rows_i64": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "spt_gather_rows_i32": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr]),
    "spt_split_tf32": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "spt_gemm_nt": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr,
                            c_i64, c_ptr]),
    "spt_gemm_tn_acc": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_i64, c_ptr,
                                c_i64, c_ptr, c_ptr]),
    "spt_segment_pool_fwd": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int, c_ptr,
                                     c_ptr, c_ptr]),
    "spt_segment_pool_bwd": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int,
                                     c_ptr, c_ptr]),
    "spt_unitsphere_workspace_bytes": (c_size, [c_i64]),
    "spt_unitsphere_fwd": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64,
                                   c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "spt_graphnorm_workspace_bytes": (c_size, [c_i64, c_i64]),
    "spt_graphnorm_fwd": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr,
                                  c_ptr, c_f32, c_f32, c_ptr, c_ptr, c_ptr, c_ptr, c_size,
                                  c_ptr]),
    "spt_graphnorm_bwd": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr,
                                  c_ptr, c_ptr, c_ptr, c_ptr, c_f32, c_ptr, c_ptr, c_ptr,
                                  c_ptr, c_ptr, c_size, c_ptr]),
    "spt_groupnorm_fwd": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_f32,
                                  c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "spt_groupnorm_bwd": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_ptr,
                                  c_ptr, c_ptr, c_f32, c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                  c_size, c_ptr]),
    "spt_attn_fwd": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr,
                             c_ptr, c_i64, c_i64, c_int, c_int, c_int, c_int, c_ptr,
                             c_ptr, c_ptr, c_ptr, c_int, c_f32, c_ptr, c_ptr, c_ptr,
                             c_ptr, c_ptr, c_ptr]),
    "spt_attn_bwd_rows": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr,  # q k v a
                                  c_ptr, c_ptr, c_i64, c_i64,                        # csr, sizes
                                  c_int, c_int, c_int, c_int,                        # H D Dv F
                                  c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_f32,          # W, scale
                                  c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,          # m z agg abar dagg dabar
                                  c_ptr, c_i64, c_ptr,                               # dq da
                                  c_ptr, c_ptr, c_ptr, c_ptr,                        # dWq dbq dWk dbk
                                  c_ptr, c_ptr, c_ptr]),                             # P G stream
    "spt_attn_fwd_ex": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr,
                                c_ptr, c_i64, c_i64, c_int, c_int, c_int, c_int, c_ptr,
                                c_ptr, c_ptr, c_ptr, c_int, c_f32, c_ptr, c_ptr, c_ptr,
                                c_ptr, c_ptr, c_ptr, c_ptr]),
    "spt_attn_bwd_rows_ex": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr,
                                     c_ptr, c_ptr, c_i64, c_i64,
                                     c_int, c_int, c_int, c_int,
                                     c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_f32,
                                     c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                     c_ptr, c_i64, c_ptr,
                                     c_ptr, c_ptr, c_ptr, c_ptr,
                                     c_ptr, c_ptr, c_ptr, c_ptr]),
    "spt_cast_bf16": (c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    "spt_attn_fwd_bf16": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr,
                                  c_ptr, c_i64, c_i64, c_int, c_int, c_int, c_int, c_ptr,
                                  c_ptr, c_ptr, c_ptr, c_int, c_f32, c_ptr, c_ptr, c_ptr,
                                  c_ptr, c_ptr, c_ptr]),
    "spt_attn_bwd_rows_bf16": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr,
                                       c_ptr, c_ptr, c_i64, c_i64,
                                       c_int, c_int, c_int, c_int,
                                       c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_f32,
                                       c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                       c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "spt_attn_bwd_targets_ex": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int, c_int, c_int,
                                        c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_ptr,
                                        c_ptr]),
    "spt_attn_bwd_targets": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_int, c_int, c_int,
                                     c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_ptr]),
    "spt_attn_bwd_weights": (c_int, [c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_ptr, c_ptr,
                                     c_ptr, c_ptr, c_ptr]),
    "spt_concat_offset_i64": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_i64, c_int, c_ptr, c_ptr]),
    "spt_relabel_consecutive_workspace_bytes": (c_size, [c_i64]),
    "spt_relabel_consecutive_i64": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr,
                                            c_ptr, c_ptr, c_size, c_ptr]),
    "spt_select_edges_workspace_bytes": (c_size, [c_i64]),
    "spt_select_edges_mark": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr,
                                      c_ptr, c_size, c_ptr]),
    "spt_select_edges_write": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "spt_csr_select_workspace_bytes": (c_size, [c_i64]),
    "spt_csr_select_pointers": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr,
                                        c_size, c_ptr]),
    "spt_csr_select_values_i64": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_ptr,
                                          c_ptr, c_ptr]),
    "spt_gather_rows_bytes": (c_int, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr]),
    "spt_sparse_sample_workspace_bytes": (c_size, [c_i64]),
    "spt_sparse_sample": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, ctypes.c_uint64,
                                  c_ptr, c_ptr, c_size, c_ptr]),
    "spt_radius_flags": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_int, ctypes.c_float, c_int, c_ptr,
                                 c_ptr, c_ptr, c_ptr]),
    "spt_khop_expand": (c_int, [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr]),
    "spt_where_workspace_bytes": (c_size, [c_i64]),
    "spt_where_count": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "spt_where_write": (c_int, [c_ptr, c_i64, c_ptr, c_ptr]),
    "spt_data_select_arena_bytes": (c_size, [c_ptr]),
    "spt_data_select": (c_int, [c_ptr, c_ptr, c_size, c_ptr, c_ptr]),
    "spt_gather_rows_multi": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_ptr, c_i64, c_ptr]),
    "spt_vrpe_blockdiag": (c_int, [c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "spt_vrpe_epilogue": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_ptr,
                                  c_ptr]),
    "spt_vrpe_bwd_params": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_int, c_ptr,
                                    c_ptr, c_ptr]),
    "spt_segment_mean_std_fwd": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr]),
    "spt_superedge_features_fwd": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr,
                                           c_ptr]),
    "spt_edge_features_fwd": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                      c_ptr, c_i64, c_i64, c_int, c_ptr, c_ptr, c_ptr]),
    "spt_vertical_edge_features_fwd": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                               c_i64, c_i64, c_ptr, c_ptr]),
}


class AttnExtras(ctypes.Structure):
    """spt_attn_extras (include/spt_b200.h): optional terms of the attention core"""
    _fields_ = [("q_row_add", ctypes.c_void_p), ("q_tgt_add", ctypes.c_void_p),
                ("k_row_add", ctypes.c_void_p), ("drop_mask", ctypes.c_void_p),
                ("d_q_row_add", ctypes.c_void_p), ("d_k_row_add", ctypes.c_void_p),
                ("d_sump", ctypes.c_void_p), ("sump", ctypes.c_void_p),
                ("ws_logits", ctypes.c_void_p), ("edge_row", ctypes.c_void_p),
                ("ws_ds", ctypes.c_void_p), ("v_bf16", ctypes.c_void_p),
                ("ldv_bf16", ctypes.c_int64)]


def library_path():
    return _build.LIB_PATH


def load():
    """Load (building first if the in-tree .so is missing or stale and nvcc is
    available).  Raises RuntimeError — never falls back to a CPU path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = _build.LIB_PATH
        if not os.path.exists(path) or _build.needs_build():
            try:
                _build.build()
            except Exception as e:  # noqa: BLE001
                raise RuntimeError(
                    f"superpoint_transformer_b200: {path} is missing and could not be "
                    f"built ({e}). The CUDA library is required; there is no fallback."
                ) from e
        try:
            lib = ctypes.CDLL(path)
        except OSError as e:
            raise RuntimeError(
                f"superpoint_transformer_b200: cannot load {path}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise RuntimeError(
                    f"superpoint_transformer_b200: {path} does not export {name}") from e
            fn.restype = res
            fn.argtypes = args
        if lib.spt_abi_version() != 1:
            raise RuntimeError("superpoint_transformer_b200: ABI version mismatch")
        _LIB = lib
    return _LIB


E_INDEX = -5   # SPT_E_INDEX


class SelectLevel(ctypes.Structure):
    """spt_select_level (include/spt_b200.h)"""
    _fields_ = [("num_nodes", c_i64), ("idx", c_ptr), ("num_selected", c_i64),
                ("edge_index", c_ptr), ("num_edges", c_i64),
                ("sub_pointers", c_ptr), ("sub_points", c_ptr),
                ("sub_items", c_i64), ("num_sub", c_i64), ("update_sub", c_int),
                ("super_index", c_ptr), ("num_super", c_i64), ("update_super", c_int),
                ("num_node_rows", c_int), ("node_src", c_ptr), ("node_row_bytes", c_ptr),
                ("num_edge_rows", c_int), ("edge_src", c_ptr), ("edge_row_bytes", c_ptr)]


SEL_NUM_EDGES, SEL_NUM_ITEMS, SEL_NUM_PARENTS, SEL_EDGE_INDEX, SEL_IDX_EDGE, SEL_SUB_POINTERS, \
    SEL_SUB_POINTS, SEL_IDX_SUB, SEL_SUB_SUPER, SEL_SUB_COUNTS, SEL_SUPER_INDEX, SEL_IDX_SUPER, \
    SEL_SUPER_SUB_POINTERS, SEL_SUPER_SUB_POINTS, SEL_ROWS = range(15)


def check(rc, what):
    if rc == E_INDEX:
        msg = load().spt_last_error()
        raise IndexError(msg.decode() if msg else what)
    if rc != 0:
        msg = load().spt_last_error()
        msg = msg.decode() if msg else ""
        raise RuntimeError(f"libspt_b200 {what} failed (status {rc}): {msg}")
