"""ctypes binding of libgae_b200.so (the C ABI declared in include/gae_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgae_b200.so")


class GaeError(RuntimeError):
    pass


class HubPlanStruct(Structure):
    """Mirror of gae_hub_plan_t."""
    _fields_ = [
        ("seg_len", c_int32),
        ("short_max", c_int32),
        ("n_long", c_int64),
        ("n_seg", c_int64),
        ("long_row", c_void_p),
        ("long_seg_ptr", c_void_p),
        ("seg_row", c_void_p),
        ("n_empty", c_int64),
        ("n_short", c_int64),
        ("n_mid", c_int64),
        ("empty_rows", c_void_p),
        ("short_rows", c_void_p),
        ("mid_rows", c_void_p),
        ("seg_order", c_void_p),
    ]


MAX_LAYERS = 8


class StepDesc(Structure):
    """Mirror of gae_step_desc_t."""
    _fields_ = [
        ("n_layers", c_int32),
        ("dims", c_int32 * (MAX_LAYERS + 1)),
        ("acts", c_int32 * MAX_LAYERS),
        ("dropout_p", c_float),
        ("pos_weight", c_float),
        ("per_graph", c_int32),
        ("x_aggregated", c_int32),
    ]


class HaloExchangeStruct(Structure):
    """Mirror of gae_halo_exchange_t."""
    _fields_ = [
        ("world", c_int32), ("rank", c_int32), ("n_stages", c_int32), ("d", c_int32),
        ("ld", c_int64),
        ("x_local", c_void_p), ("peer_x", c_void_p), ("peer_flags", c_void_p), ("flags", c_void_p),
        ("send_src", c_void_p), ("send_peer", c_void_p), ("send_dst", c_void_p),
        ("stage_ptr", c_void_p), ("stage_done", c_void_p),
        ("push_ctas", c_int32), ("push_threads", c_int32), ("timeout_ms", c_int32),
        ("pre_rowptr", c_void_p), ("pre_col", c_void_p), ("pre_plan", c_void_p), ("pre_ws", c_void_p),
        ("pre_n_rows", c_int64), ("pre_row0", c_int64), ("pre_stage", c_int32),
    ]


class HaloBlockStruct(Structure):
    """Mirror of gae_halo_block_t."""
    _fields_ = [
        ("row0", c_int64), ("n_rows", c_int64),
        ("rowptr", c_void_p), ("col", c_void_p), ("plan", c_void_p), ("partial_ws", c_void_p),
        ("accumulate", c_int32),
    ]


# name -> (restype, argtypes); every symbol include/gae_b200.h declares
SIGNATURES = {
    "gae_version": (c_char_p, []),
    "gae_last_error_string": (c_char_p, []),
    "gae_set_tuning": (c_int, [c_char_p, c_int32]),
    "gae_get_tuning": (c_int32, [c_char_p]),
    "gae_launch_count": (c_int64, []),
    "gae_hub_plan_count_host": (c_int, [c_void_p, c_int64, c_int32, POINTER(c_int64), POINTER(c_int64)]),
    "gae_hub_plan_fill_host": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "gae_row_bins_host": (c_int, [c_void_p, c_int64, c_int32, c_int32, POINTER(c_int64 * 3), c_void_p, c_void_p,
                                  c_void_p]),
    "gae_spmm_csr_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                 c_int32, POINTER(HubPlanStruct), c_void_p, c_int32, c_void_p]),
    "gae_spmm_csr_f32_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
                                      c_int32, POINTER(HubPlanStruct), c_void_p, c_void_p, c_void_p, c_void_p]),
    "gae_linear_fwd_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
                                   c_int32, c_int32, c_void_p]),
    "gae_linear_bwd_ws_bytes": (c_int64, [c_int64, c_int32, c_int32]),
    "gae_linear_bwd_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                   c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32,
                                   c_int32, c_void_p]),
    "gae_gcn_layer_fwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                      c_void_p, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p]),
    "gae_gcn_layer_bwd_ws_bytes": (c_int64, [c_int64, c_int32, c_int32]),
    "gae_gcn_layer_bwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                      c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                      c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p]),
    "gae_dropout_fwd_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_float,
                                    c_uint64, c_uint64, c_int32, c_void_p]),
    "gae_dropout_fwd_devrng_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_float,
                                           c_void_p, c_void_p]),
    "gae_dropout_bwd_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_float,
                                    c_void_p, c_void_p]),
    "gae_decoder_ws_bytes": (c_int64, [c_int64, c_int32]),
    "gae_decoder_bce_f32": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_float, c_int32, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "gae_decoder_blockdiag_ws_bytes": (c_int64, [c_int64, c_int32]),
    "gae_decoder_bce_blockdiag_f32": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_double, c_float, c_int32, c_void_p, c_void_p,
                                              c_int64, c_void_p, c_int64, c_void_p]),
    "gae_decoder_logits_f32": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int64, c_void_p]),
    "gae_step_ws_bytes": (c_int64, [POINTER(StepDesc), c_int64, POINTER(HubPlanStruct), POINTER(HubPlanStruct)]),
    "gae_step_fwd_bwd_f32": (c_int, [POINTER(StepDesc), c_int64, c_void_p, c_void_p, POINTER(HubPlanStruct), c_void_p,
                                     c_void_p, POINTER(HubPlanStruct), c_void_p, c_int64, POINTER(c_void_p),
                                     POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int32,
                                     c_void_p, c_void_p, c_int64, POINTER(c_void_p), POINTER(c_void_p), c_void_p,
                                     c_int64, c_void_p]),
    "gae_in_degrees_i64": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "gae_batch_offset_cols_i32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "gae_batch_assemble": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                   c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p]),
    "gae_gather_rows_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p]),
    "gae_pull_rows_p2p_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int64,
                                      c_void_p]),
    "gae_push_rows_p2p_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                      c_int32, c_void_p]),
    "gae_ipc_get_handle": (c_int, [c_void_p, POINTER(c_uint8 * 64), POINTER(c_int64)]),
    "gae_ipc_open_handle": (c_int, [POINTER(c_uint8 * 64), POINTER(c_void_p)]),
    "gae_ipc_close_handle": (c_int, [c_void_p]),
    "gae_decoder_tile_probe_f32": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                           c_void_p, POINTER(c_int32), c_void_p]),
    "gae_halo_push_f32": (c_int, [POINTER(HaloExchangeStruct), c_uint64, c_void_p]),
    "gae_halo_push_range_f32": (c_int, [POINTER(HaloExchangeStruct), c_uint64, c_int32, c_int32, c_void_p]),
    "gae_halo_wait_f32": (c_int, [POINTER(HaloExchangeStruct), c_int32, c_uint64, c_void_p]),
    "gae_halo_release_f32": (c_int, [POINTER(HaloExchangeStruct), c_uint64, c_void_p]),
    "gae_halo_spmm_f32": (c_int, [POINTER(HaloExchangeStruct), POINTER(HaloBlockStruct), c_void_p, c_int64, c_uint64,
                                  c_void_p, c_void_p, c_void_p]),
    "gae_halo_status": (c_int, [POINTER(HaloExchangeStruct), POINTER(c_int64)]),
    "gae_halo_plan_count_host": (c_int, [c_void_p, c_int64, c_void_p, c_int32, c_int32, POINTER(c_int64), c_void_p]),
    "gae_halo_plan_fill_host": (c_int, [c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "gae_halo_stage_tags_host": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int32, c_void_p]),
    "gae_halo_push_lists_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p,
                                         c_void_p, c_void_p, c_void_p]),
    "gae_adam_step_f32": (c_int, [c_int32, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                  POINTER(c_int64), c_float, c_float, c_float, c_float, c_int64, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library and bind every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GaeError(
            f"{LIB_PATH} is missing: the CUDA library has not been built. "
            "Run `python -m gae_dgl_b200.build` (needs nvcc). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gae_last_error_string().decode(errors="replace")
        raise GaeError(f"{what} failed with code {rc}: {msg}")


def set_tuning(key: str, value: int) -> None:
    check(load().gae_set_tuning(key.encode(), int(value)), "gae_set_tuning")


def get_tuning(key: str) -> int:
    return int(load().gae_get_tuning(key.encode()))


def launch_count() -> int:
    return int(load().gae_launch_count())


def version() -> str:
    return load().gae_version().decode()
