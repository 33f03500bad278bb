"""ctypes loader for libcp3b200.so - the C ABI declared in include/cp3b200.h.

This is the same stub a maintainer of the reference would write to bind the drop-in.
There is no CPU fallback: if the shared library is missing the import fails loudly, and every call that
needs the device aborts / returns an error when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libcp3b200.so")


class cp3g_hit(C.Structure):
    _fields_ = [("req", C.c_uint32), ("clause", C.c_uint32), ("lit", C.c_int32)]


CP_SYMBOLS = [
    "CPinit", "CPpreprocess", "CPdestroy", "CPparseOptions", "CPparseOptionString", "CPsetPresetConfig",
    "CPaddLit", "CPversion", "CPhasNextOutputLit", "CPnextOutputLit", "CPresetOutput", "CPfreezeVariable",
    "CPgetReplaceLiteral", "CPresetModel", "CPpushModelBool", "CPgiveModelLit", "CPpostprocessModel",
    "CPgetFinalModelLit", "CPmodelVariables", "CPnVars", "CPnClss", "CPnLits", "CPok", "CPlitFalsified",
    "CPaddLits", "CPgetOutput", "CPoutputUnits", "CPgetUndo", "CPstat", "CPsetDistributed",
]
CP3G_SYMBOLS = [
    "cp3g_device_count", "cp3g_create", "cp3g_destroy", "cp3g_last_error", "cp3g_upload", "cp3g_upload_packed", "cp3g_build",
    "cp3g_download_occ", "cp3g_set_deleted", "cp3g_shrink", "cp3g_append", "cp3g_scan", "cp3g_scan_all", "cp3g_bve_pairs", "cp3g_bce_masks",
    "cp3g_bve_resolve", "cp3g_sub_pass", "cp3g_sub_pass_fetch", "cp3g_str_static", "cp3g_bve_eval", "cp3g_bve_stream", "cp3g_dense", "cp3g_compact", "cp3g_ingest", "cp3g_download_arena", "cp3g_nccl_unique_id", "cp3g_nccl_init", "cp3g_counter",
]


def load(path: str | None = None) -> C.CDLL:
    path = path or os.environ.get("CP3B200_LIB") or DEFAULT_LIB
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `make -C riss-solver_b200/csrc` (or __graft_entry__.build()). "
            "The Coprocessor GPU path has no CPU fallback.")
    L = C.CDLL(path)
    vp, i32p, u32p = C.c_void_p, C.c_void_p, C.c_void_p
    L.CPinit.restype = vp; L.CPinit.argtypes = [C.c_char_p]
    L.CPpreprocess.argtypes = [vp]; L.CPpreprocess.restype = None
    L.CPdestroy.argtypes = [C.POINTER(vp)]; L.CPdestroy.restype = None
    L.CPparseOptions.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_char_p), C.c_int]; L.CPparseOptions.restype = None
    L.CPparseOptionString.argtypes = [vp, C.c_char_p]; L.CPparseOptionString.restype = None
    L.CPsetPresetConfig.argtypes = [vp, C.c_char_p]; L.CPsetPresetConfig.restype = None
    L.CPaddLit.argtypes = [vp, C.c_int]; L.CPaddLit.restype = None
    L.CPversion.argtypes = [vp]; L.CPversion.restype = C.c_float
    for f in ("CPhasNextOutputLit", "CPnextOutputLit", "CPgetFinalModelLit", "CPmodelVariables", "CPnVars",
              "CPnClss", "CPnLits", "CPok"):
        getattr(L, f).argtypes = [vp]; getattr(L, f).restype = C.c_int
    for f in ("CPresetOutput", "CPresetModel", "CPpostprocessModel"):
        getattr(L, f).argtypes = [vp]; getattr(L, f).restype = None
    L.CPfreezeVariable.argtypes = [vp, C.c_int]; L.CPfreezeVariable.restype = None
    L.CPgetReplaceLiteral.argtypes = [vp, C.c_int]; L.CPgetReplaceLiteral.restype = C.c_int
    L.CPpushModelBool.argtypes = [vp, C.c_int]; L.CPpushModelBool.restype = None
    L.CPgiveModelLit.argtypes = [vp, C.c_int]; L.CPgiveModelLit.restype = None
    L.CPlitFalsified.argtypes = [vp, C.c_int]; L.CPlitFalsified.restype = C.c_int
    L.CPaddLits.argtypes = [vp, i32p, C.c_size_t]; L.CPaddLits.restype = None
    L.CPgetOutput.argtypes = [vp, i32p, C.c_size_t]; L.CPgetOutput.restype = C.c_size_t
    L.CPoutputUnits.argtypes = [vp]; L.CPoutputUnits.restype = C.c_size_t
    L.CPgetUndo.argtypes = [vp, i32p, C.c_size_t]; L.CPgetUndo.restype = C.c_size_t
    L.CPstat.argtypes = [vp, C.c_char_p]; L.CPstat.restype = C.c_int64
    L.CPsetDistributed.argtypes = [vp, C.c_int, C.c_int, vp]; L.CPsetDistributed.restype = C.c_int
    # pass-level scan service
    L.cp3g_device_count.restype = C.c_int
    L.cp3g_create.argtypes = [C.c_int]; L.cp3g_create.restype = vp
    L.cp3g_destroy.argtypes = [vp]; L.cp3g_destroy.restype = None
    L.cp3g_last_error.argtypes = [vp]; L.cp3g_last_error.restype = C.c_char_p
    L.cp3g_upload.argtypes = [vp, C.c_uint32, C.c_uint32, i32p, C.c_size_t, u32p, u32p, vp]; L.cp3g_upload.restype = C.c_int
    L.cp3g_build.argtypes = [vp]; L.cp3g_build.restype = C.c_int
    L.cp3g_download_occ.argtypes = [vp, u32p, u32p]; L.cp3g_download_occ.restype = C.c_int
    L.cp3g_set_deleted.argtypes = [vp, C.c_uint32, u32p]; L.cp3g_set_deleted.restype = C.c_int
    L.cp3g_shrink.argtypes = [vp, C.c_uint32, u32p, u32p, u32p, i32p, C.c_size_t]; L.cp3g_shrink.restype = C.c_int
    L.cp3g_append.argtypes = [vp, C.c_uint32, u32p, i32p, C.c_size_t]; L.cp3g_append.restype = C.c_int
    L.cp3g_scan.argtypes = [vp, C.c_int, C.c_uint32, u32p, u32p, i32p, vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    L.cp3g_scan.restype = C.c_int
    L.cp3g_scan_all.argtypes = [vp, C.c_int, vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    L.cp3g_scan_all.restype = C.c_int
    L.cp3g_bve_pairs.argtypes = [vp, C.c_uint32, u32p, u32p, u32p, u32p, u32p, vp, vp]; L.cp3g_bve_pairs.restype = C.c_int
    if hasattr(L, "cp3g_bce_masks"):      # (absent only in builds older than the header; tests/test_abi.py checks every declared symbol)
        L.cp3g_bce_masks.argtypes = [vp, C.c_uint32, u32p, u32p, u32p, u32p, u32p, vp, u32p]; L.cp3g_bce_masks.restype = C.c_int
    L.cp3g_bve_resolve.argtypes = [vp, C.c_uint32, u32p, u32p, u32p, vp, i32p, C.c_size_t]; L.cp3g_bve_resolve.restype = C.c_int
    L.cp3g_sub_pass.argtypes = [vp, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]; L.cp3g_sub_pass.restype = C.c_int
    if hasattr(L, "cp3g_str_static"):
        L.cp3g_str_static.argtypes = [vp, u32p, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), u32p]; L.cp3g_str_static.restype = C.c_int
    L.cp3g_sub_pass_fetch.argtypes = [vp, u32p, u32p, u32p, u32p]; L.cp3g_sub_pass_fetch.restype = C.c_int
    L.cp3g_bve_eval.argtypes = [vp, C.c_uint32, u32p, u32p, u32p, u32p, u32p, C.c_int, vp, vp, vp]; L.cp3g_bve_eval.restype = C.c_int
    L.cp3g_bve_stream.argtypes = [vp, C.c_uint32, u32p, vp, vp, vp, i32p]; L.cp3g_bve_stream.restype = C.c_int
    L.cp3g_dense.argtypes = [vp, vp, C.c_int, u32p, u32p, C.POINTER(C.c_int), i32p, C.c_size_t]; L.cp3g_dense.restype = C.c_int
    L.cp3g_ingest.argtypes = [vp, i32p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]; L.cp3g_ingest.restype = C.c_int
    L.cp3g_download_arena.argtypes = [vp, i32p, vp]; L.cp3g_download_arena.restype = C.c_int
    L.cp3g_compact.argtypes = [vp, u32p, i32p, C.c_size_t, C.POINTER(C.c_size_t)]; L.cp3g_compact.restype = C.c_int
    L.cp3g_nccl_unique_id.argtypes = [vp]; L.cp3g_nccl_unique_id.restype = C.c_int
    L.cp3g_nccl_init.argtypes = [vp, C.c_int, C.c_int, vp]; L.cp3g_nccl_init.restype = C.c_int
    L.cp3g_counter.argtypes = [vp, C.c_char_p]; L.cp3g_counter.restype = C.c_int64
    return L
