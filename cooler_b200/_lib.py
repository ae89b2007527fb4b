"""ctypes binding of the C-ABI library ``libcooler_b200.so`` (include/cooler_b200.h).

There is NO fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CB200_LIB selects another build of the same ABI (e.g. a far-field row-block size variant made by
# `python -m cooler_b200.build --far-row-shift 13`); default: the in-tree library.
LIB_PATH = os.environ.get("CB200_LIB") or os.path.join(_HERE, "libcooler_b200.so")

KEEP_ALL, KEEP_CIS, KEEP_TRANS, KEEP_BAND_NEAR, KEEP_BAND_FAR, KEEP_EARLIER_CHROM = 0, 1, 2, 3, 4, 5


class Copy(C.Structure):
    """cb200_copy"""
    _fields_ = [("px", C.c_void_p), ("indptr", C.c_void_p), ("tile_row", C.c_void_p),
                ("nnz", C.c_int64)]


class Sub(C.Structure):
    """cb200_sub"""
    _fields_ = [("px", C.c_void_p), ("indptr", C.c_void_p), ("tile_row", C.c_void_p),
                ("nnz", C.c_int64), ("direct", C.c_void_p), ("pieces", C.c_void_p),
                ("first_warp", C.c_int64), ("first_tile", C.c_int64),
                ("gather_base", C.c_int32), ("gather_len", C.c_int32),
                ("row_base", C.c_int32), ("row_count", C.c_int32), ("val", C.c_void_p)]


class Ice(C.Structure):
    """cb200_ice"""
    _fields_ = [
        ("subs", C.c_void_p), ("n_subs", C.c_int32), ("span", C.c_int32),
        ("total_warps", C.c_int64), ("total_tiles", C.c_int64), ("smem_bins", C.c_int32),
        ("claim", C.c_void_p),
        ("n_bins", C.c_int32), ("n_segs", C.c_int32),
        ("seg_offset", C.c_void_p),
        ("n_blocks", C.c_int32),
        ("blk_seg", C.c_void_p), ("blk_lo", C.c_void_p), ("blk_hi", C.c_void_p),
        ("seg_blk_offset", C.c_void_p),
        ("bias", C.c_void_p), ("cweights", C.c_void_p), ("w", C.c_void_p),
        ("s", C.c_void_p), ("marg", C.c_void_p),
        ("blk_sum", C.c_void_p), ("blk_cnt", C.c_void_p), ("blk_dev", C.c_void_p),
        ("seg_mean", C.c_void_p), ("seg_nnz", C.c_void_p),
        ("seg_scale", C.c_void_p), ("seg_var", C.c_void_p),
        ("seg_iters", C.c_void_p), ("seg_done", C.c_void_p), ("seg_empty", C.c_void_p),
        ("all_done", C.c_void_p), ("tickets", C.c_void_p), ("var_hist", C.c_void_p),
        ("tol", C.c_double), ("max_iters", C.c_int32),
        ("n_far_copies", C.c_int32),
        ("far_copies", C.c_void_p), ("far_units", C.c_void_p), ("far_tiles", C.c_void_p),
        ("far_parts", C.c_void_p), ("far_claim", C.c_void_p),
        ("n_far_units", C.c_int32), ("far_col_shift", C.c_int32), ("far_warps", C.c_int32),
        ("far_rec_bytes", C.c_int32), ("float_counts", C.c_int32),
    ]


class FarCopy(C.Structure):
    """cb200_far_copy"""
    _fields_ = [("px", C.c_void_p), ("ptr", C.c_void_p), ("nnz", C.c_int64),
                ("rb0", C.c_int32), ("n_rb", C.c_int32), ("j0", C.c_int32), ("n_j", C.c_int32),
                ("rb_unit", C.c_void_p)]


class Remap(C.Structure):
    """cb200_remap"""
    _fields_ = [("bin_map", C.c_void_p), ("blocks", C.c_void_p), ("block_shift", C.c_int32),
                ("magic", C.c_uint32), ("magic_shift", C.c_int32)]


class Chunk(C.Structure):
    """cb200_chunk"""
    _fields_ = [("src", C.c_void_p), ("src_bytes", C.c_int64), ("dst_row", C.c_int64),
                ("row0", C.c_int32), ("row1", C.c_int32), ("filter_mask", C.c_uint32)]


class FarUnit(C.Structure):
    """cb200_far_unit"""
    _fields_ = [("copy", C.c_int32), ("rb", C.c_int32), ("t0", C.c_int32), ("t1", C.c_int32)]


def far_row_shift():
    """log2 rows per far-field block of the loaded library (CB200_FAR_ROW_SHIFT, build-time)."""
    return int(load().cb200_far_row_shift())


_P, _I32, _I64, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/cooler_b200.h
SIGNATURES = {
    "cb200_abi_version": (C.c_int, []),
    "cb200_last_error": (C.c_char_p, []),
    "cb200_set_device": (C.c_int, [C.c_int]),
    "cb200_warp_tile": (C.c_int, []),
    "cb200_scan_workspace_bytes": (_SZ, [_I64]),
    "cb200_exclusive_scan_i32": (C.c_int, [_P, _P, _I64, _P, _SZ, _P]),
    "cb200_sort_workspace_bytes": (_SZ, [_I64]),
    "cb200_sort_pairs": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, C.c_int, C.c_int, _P, _SZ, _P]),
    "cb200_median_workspace_bytes": (_SZ, [_I32]),
    "cb200_segment_medians": (C.c_int, [_P, _P, _I32, _I64, _I32, _P, _P, _P, _SZ, _P]),
    "cb200_pack_pixels": (C.c_int, [_P, _P, _I64, _P, _P]),
    "cb200_pack_pixels_i32": (C.c_int, [_P, _P, _I64, _P, _P]),
    "cb200_tile_rows": (C.c_int, [_P, _I32, _I64, _P, _P]),
    "cb200_indptr_from_sorted": (C.c_int, [_P, _I64, _I32, _I32, _I32, _P, _P]),
    "cb200_swap_key_other": (C.c_int, [_P, _P, _I64, _P, _P, _P]),
    "cb200_strip_finish": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P, _P]),
    "cb200_far_finish": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "cb200_far_slot_chunks": (C.c_int, [_P, _I64, _P, _P]),
    "cb200_far_pieces": (C.c_int, [_P, _I64, _P, _P, _P]),
    "cb200_far_expand": (C.c_int, [_P, _P, _I64, _P, _P, _P, _P]),
    "cb200_far_finish4": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "cb200_far_interleave4": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _P, _P]),
    "cb200_far_interleave": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _P, _P]),
    "cb200_filter_count": (C.c_int, [_P, _P, _P, _I32, _I64, _P, _P, _I32, _I32, _I32, _I32, _P, _P]),
    "cb200_filter_write": (C.c_int, [_P, _P, _P, _I32, _I64, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "cb200_marginals_i64": (C.c_int, [C.POINTER(Copy), _I32, _P, _P, _P, _P]),
    "cb200_attach_values": (C.c_int, [_P, _I64, _P, _I64, _P, _P]),
    "cb200_values_nonzero": (C.c_int, [_P, _I64, _P, _P]),
    "cb200_cis_touch": (C.c_int, [_P, _P, _I64, _P, _I32, _P, _P, _P, _P]),
    "cb200_set_tuning": (C.c_int, [_I32, _I32]),
    "cb200_far_row_shift": (C.c_int, []),
    "cb200_ice_sweep": (C.c_int, [C.POINTER(Ice), _P]),
    "cb200_ice_combine": (C.c_int, [C.POINTER(Ice), _P]),
    "cb200_ice_update": (C.c_int, [C.POINTER(Ice), _I32, _I32, _P]),
    "cb200_ice_update_peer": (C.c_int, [C.POINTER(Ice), _I32, _P, _I32, _I64, _P]),
    "cb200_ice_reduce_peer": (C.c_int, [C.POINTER(Ice), _P, _I32, _I32, _I64, _P, _P]),
    "cb200_ice_update_gather": (C.c_int, [C.POINTER(Ice), _I32, _P, _I32, _I64, _P]),
    "cb200_decode_chunks": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _I32]),
    "cb200_bin_pairs": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P]),
    "cb200_fetch_count": (C.c_int, [C.POINTER(Copy), _I32, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "cb200_fetch_dense": (C.c_int, [C.POINTER(Copy), _I32, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "cb200_fetch_coo": (C.c_int, [C.POINTER(Copy), _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "cb200_fetch_balance_dense": (C.c_int, [_P, _I64, _I64, _P, _P, _I32, _P, _P]),
    "cb200_fetch_balance_coo": (C.c_int, [_P, _P, _P, _I64, _P, _P, _I32, _P, _P]),
    "cb200_fetch_dense_f64": (C.c_int, [C.POINTER(Copy), _P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "cb200_fetch_coo_f64": (C.c_int, [C.POINTER(Copy), _P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "cb200_fetch_balance_dense_f64": (C.c_int, [_P, _I64, _I64, _P, _P, _I32, _P, _P]),
    "cb200_fetch_balance_coo_f64": (C.c_int, [_P, _P, _P, _I64, _P, _P, _I32, _P, _P]),
    "cb200_coarsen_remap": (C.c_int, [C.POINTER(Copy), _I32, _P, _P, _P, _P]),
    "cb200_coarsen_maps": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "cb200_coarsen_block_table": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P]),
    "cb200_coarsen_merge_round": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, C.POINTER(Remap), _P, _P, _P]),
    "cb200_coarsen_merge_reduce": (C.c_int, [_P, _P, _P, _P, _I32, _I32, C.POINTER(Remap), _P, _P, _P, _P]),
    "cb200_coarsen_compact": (C.c_int, [_P, _P, _P, _P, _I32, _P, _P, _P]),
    "cb200_runs_count": (C.c_int, [_P, _P, _I64, _P, _P]),
    "cb200_runs_write": (C.c_int, [_P, _P, _I64, _P, _I32, _P, _P, _P, _P, _P]),
    "cb200_runs_aggregate": (C.c_int, [_P, _P, _I64, _P, _P, _I32, _I32, _P, _P, _P, _P, _P]),
    "cb200_unpack_pixels": (C.c_int, [_P, _P, _I64, _P, _P, _P, _P]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Load the library (once).  Raises NativeLibraryError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m cooler_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().cb200_last_error().decode("utf-8", "replace")
        raise NativeLibraryError(f"{what} failed (rc={rc}): {msg}")


def call(name, *args):
    check(getattr(load(), name)(*args), name)
