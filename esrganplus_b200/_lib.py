"""ctypes binding of libesrp.so (C ABI declared in include/esrp.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or a call
fails, a RuntimeError is raised.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C esrganplus_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

ESRP_MAX_CHUNKS = 32
LAYOUT_TILE = 0   # kx-stacked RM x CW tiles (conv3x3_tc.cuh): any width, narrow images
LAYOUT_ROW = 1    # ky-stacked row streaming (conv3x3_row.cuh): images wider than ~64 px, bn <= 32


VARIANT_ROW_ALT = 0x2000   # ESRP_VARIANT_ROW_ALT: row kernel, MMA issuers alternate whole rows (include/esrp.h)
VARIANT_PAIR = 0x4000      # ESRP_VARIANT_PAIR: row kernel, clusters of two CTAs share every MMA (cta_group::2)


def variant_mt(mt: int) -> int:
    """esrp_conv3x3_t.variant bits forcing `mt` accumulator slots per CTA tile (ESRP_VARIANT_MT)."""
    return mt & 15


def variant_cwlog2(l: int) -> int:
    """variant bits forcing the M-tile width 2**l pixels (ESRP_VARIANT_CWLOG2)."""
    return (l & 15) << 4

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ESRP_LIBRARY") or os.path.join(_HERE, "libesrp.so")   # (ESRP_LIBRARY: A/B runs of two builds)


class Conv3x3Desc(C.Structure):
    """Mirror of ``esrp_conv3x3_t`` (include/esrp.h) — field order and types must match."""

    _fields_ = [
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("src", C.c_void_p * 2),
        ("src_ctotal", C.c_int32 * 2),
        ("kc", C.c_int32),
        ("num_chunks", C.c_int32),
        ("chunk_src", C.c_int32 * ESRP_MAX_CHUNKS),
        ("chunk_c0", C.c_int32 * ESRP_MAX_CHUNKS),
        ("aux_chunks", C.c_int32),
        ("bn", C.c_int32),
        ("cout", C.c_int32),
        ("w_packed", C.c_void_p),
        ("w_layout", C.c_int32),
        ("bias", C.c_void_p),
        ("act", C.c_int32),
        ("s0", C.c_float),
        ("r1", C.c_void_p),
        ("r1_is_f32", C.c_int32), ("r1_ctotal", C.c_int32), ("r1_c0", C.c_int32),
        ("s1", C.c_float),
        ("r2", C.c_void_p),
        ("r2_is_f32", C.c_int32), ("r2_ctotal", C.c_int32), ("r2_c0", C.c_int32),
        ("s2", C.c_float),
        ("noise", C.c_int32),
        ("noise_ctotal", C.c_int32), ("noise_c0", C.c_int32),
        ("sigma", C.c_float),
        ("seed", C.c_uint64), ("offset", C.c_uint64),
        ("out_bf16", C.c_void_p),
        ("ob_ctotal", C.c_int32), ("ob_c0", C.c_int32),
        ("out_f32", C.c_void_p),
        ("of_ctotal", C.c_int32), ("of_c0", C.c_int32),
        ("out_nchw", C.c_void_p),
        ("variant", C.c_int32),
        ("trace", C.c_void_p),
        # training extensions
        ("mask_out", C.c_void_p),
        ("mask_out_ctotal", C.c_int32), ("mask_out_c0", C.c_int32),
        ("mask_in", C.c_void_p),
        ("mask_in_ctotal", C.c_int32), ("mask_in_c0", C.c_int32),
        ("r2_pre", C.c_int32),
        ("pre_bf16", C.c_void_p),
        ("pb_ctotal", C.c_int32), ("pb_c0", C.c_int32),
        ("pre_f32", C.c_void_p),
        ("pf_ctotal", C.c_int32), ("pf_c0", C.c_int32),
        # co-scheduled output slices (row kernel)
        ("slices", C.c_int32),
        ("slice_stride", C.c_int64),
        ("f32_planar", C.c_int32),
        ("k_valid", C.c_int32),
        ("out_lo", C.c_int32),
        ("ob_lo_c0", C.c_int32),
    ]


WGRAD_MAX_UNITS = 256


class DgradGroup(C.Structure):
    """Mirror of ``esrp_dgrad_group_t``."""
    _fields_ = [("w", C.c_void_p), ("w_o", C.c_int32), ("w_i", C.c_int32), ("co0", C.c_int32), ("scale", C.c_float)]


class WgradUnit(C.Structure):
    """Mirror of ``esrp_wgrad_unit_t``."""
    _fields_ = [("x", C.c_void_p), ("x_ctotal", C.c_int32), ("x_c0", C.c_int32),
                ("dy", C.c_void_p), ("dy_ctotal", C.c_int32), ("dy_c0", C.c_int32),
                ("acc", C.c_void_p), ("bias_acc", C.c_void_p)]


class ScatterEntry(C.Structure):
    """Mirror of ``esrp_scatter_entry_t``."""
    _fields_ = [("acc", C.c_void_p), ("dst", C.c_void_p), ("dst_off", C.c_int64), ("dst_index", C.c_int32),
                ("kind", C.c_int32), ("col0", C.c_int32), ("ncols", C.c_int32), ("nci", C.c_int32),
                ("co0", C.c_int32), ("ci0", C.c_int32), ("w_i", C.c_int32), ("scale", C.c_float)]


# name -> (restype, argtypes); the single source of truth for the symbols tests check.
SYMBOLS = {
    "esrp_last_error": (C.c_char_p, []),
    "esrp_version": (C.c_int, []),
    "esrp_sm_count": (C.c_int, []),
    "esrp_sizeof_conv3x3": (C.c_int32, []),
    "esrp_philox_normal_host": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p]),
    "esrp_conv3x3_nhwc": (C.c_int, [C.POINTER(Conv3x3Desc), C.c_void_p]),
    "esrp_packed_conv3x3_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "esrp_pack_conv3x3_weights": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_void_p,
                                            C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "esrp_nchw_f32_to_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_void_p]),
    "esrp_nhwc_bf16_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_void_p]),
    "esrp_upsample2x_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_void_p]),
    "esrp_s2d_pad_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "esrp_bn_stats_nhwc_f32": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p]),
    "esrp_bn_apply_nhwc": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esrp_streams_create": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p)]),
    "esrp_streams_fork": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32]),
    "esrp_streams_join": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32]),
    "esrp_linear_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_void_p]),
    "esrp_bn_bwd_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esrp_bn_bwd_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esrp_bn_finalize": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int32,
                                   C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "esrp_bn_bwd_finalize": (C.c_int, [C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "esrp_s2d_pad_bwd_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_void_p]),
    "esrp_linear_bwd_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "esrp_pack_dgrad_weights": (C.c_int, [C.POINTER(DgradGroup), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_void_p, C.c_void_p]),
    "esrp_conv3x3_wgrad": (C.c_int, [C.POINTER(WgradUnit), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p]),
    "esrp_wgrad_scatter": (C.c_int, [C.POINTER(ScatterEntry), C.c_int32, C.c_void_p]),
    "esrp_conv1x1_bwd": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "esrp_upsample2x_bwd_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esrp_rrdbnet_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_void_p)]),
    "esrp_rrdbnet_destroy": (None, [C.c_void_p]),
    "esrp_rrdbnet_num_tensors": (C.c_int32, [C.c_void_p]),
    "esrp_rrdbnet_tensor_key": (C.c_char_p, [C.c_void_p, C.c_int32]),
    "esrp_rrdbnet_tensor_shape": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "esrp_rrdbnet_load_weights": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]),
    "esrp_rrdbnet_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "esrp_rrdbnet_num_launches": (C.c_int32, [C.c_void_p]),
    "esrp_rrdbnet_set_chain": (C.c_int, [C.c_void_p, C.c_int32]),
    "esrp_rrdbnet_num_chained_convs": (C.c_int32, [C.c_void_p]),
    "esrp_rrdbnet_num_pair_launches": (C.c_int32, [C.c_void_p]),
    "esrp_rrdbnet_set_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "esrp_rrdbnet_get_timing": (C.c_int32, [C.c_void_p, C.POINTER(C.c_float), C.c_int32]),
    "esrp_rrdbnet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_void_p]),
    "esrp_u8hwc_to_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_void_p]),
    "esrp_nchw_f32_to_u8hwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_void_p]),
    "esrp_rrdbnet_workspace_bytes_u8": (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "esrp_rrdbnet_forward_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "esrp_rrdbnet_train_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "esrp_rrdbnet_train_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_void_p]),
    "esrp_rrdbnet_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p,
                                        C.c_void_p]),
    "esrp_rrdbnet_train_num_launches": (C.c_int32, [C.c_void_p, C.c_int32]),
    "esrp_adam_flat": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double,
                                 C.c_double, C.c_double, C.c_int64, C.c_void_p]),
    "esrp_ragan_bce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "esrp_l1_loss_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esrp_lrhr_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esrp_sizeof_lrhr_job": (C.c_int32, []),
    "esrp_graph_begin": (C.c_int, [C.c_void_p]),
    "esrp_graph_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "esrp_graph_abort": (C.c_int, [C.c_void_p]),
    "esrp_graph_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "esrp_graph_destroy": (None, [C.c_void_p]),
    "esrp_memset_zero": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "esrp_nchw_f32_to_nhwc_bf16_affine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                    C.c_int32, C.c_int32, C.c_void_p]),
    "esrp_maxpool2x2_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "esrp_relu_bwd_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "esrp_maxpool2x2_relu_bwd_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                     C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load libesrp.so once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "There is no CPU/PyTorch fallback for the hot path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().esrp_last_error()
        raise RuntimeError(f"libesrp {what} failed: {msg.decode() if msg else rc}")
