"""ctypes view of libkhg_b200.so (include/khg_b200.h).

This is the raw boundary: plain pointers and sizes.  Buffers may be numpy arrays
(host) or torch CUDA tensors (device) — the location flag is derived from the
object.  There is no CPU fallback: if the shared library is missing, or no CUDA
device is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# KHG_B200_LIB: development override used for same-box A/B timing of two builds (tools/ab_build.sh)
LIB_PATH = os.environ.get("KHG_B200_LIB") or os.path.join(_HERE, "libkhg_b200.so")

KHG_OK, KHG_ERR_INVALID, KHG_ERR_CUDA, KHG_ERR_NONFINITE, KHG_ERR_UNSUPPORTED = range(5)
KHG_HOST, KHG_DEVICE = 0, 1
KHG_FRAME_MAJOR, KHG_PDF_MAJOR = 0, 1
KHG_KERNEL_AUTO, KHG_KERNEL_SIMT, KHG_KERNEL_TCGEN05, KHG_KERNEL_TCGEN05_F16, KHG_KERNEL_TCGEN05_F16_GS = 0, 1, 2, 3, 4

# Every symbol include/khg_b200.h declares: (name, restype, argtypes)
_vp = C.c_void_p
_i32, _i64, _u16, _f32 = C.c_int32, C.c_int64, C.c_uint16, C.c_float
SYMBOLS = [
    ("khg_last_error", C.c_char_p, []),
    ("khg_abi_version", _i32, []),
    ("khg_device_count", _i32, [C.POINTER(_i32)]),
    ("khg_set_device", _i32, [_i32]),
    ("khg_augment_flags", _u16, [_u16]),
    ("khg_model_create", _i32, [_i32, _i32, _vp, C.POINTER(_vp)]),
    ("khg_model_upload", _i32, [_vp, _vp, _vp, _vp, _vp, C.POINTER(_i32)]),
    ("khg_model_get_gconsts", _i32, [_vp, _vp]),
    ("khg_model_info", _i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    ("khg_model_dense_kernel", _i32, [_vp, C.POINTER(_i32)]),
    ("khg_model_stats_kernel", _i32, [_vp, C.POINTER(_i32)]),
    ("khg_model_set_kernel", _i32, [_vp, _i32]),
    ("khg_model_set_stream", _i32, [_vp, _vp]),
    ("khg_model_sync", _i32, [_vp]),
    ("khg_model_destroy", None, [_vp]),
    ("khg_compute_gconsts", _i32, [_i32, _i32, _vp, _vp, _vp, _vp, C.POINTER(_i32)]),
    ("khg_loglikes_all_pdfs", _i32, [_vp, _vp, _i64, _i32, _f32, _i32, _vp, _i64, _i32]),
    ("khg_loglikes_pdf_subset", _i32, [_vp, _vp, _i64, _i32, _vp, _i32, _f32, _vp, _i64, _i32]),
    ("khg_pdf_loglikes", _i32, [_vp, _i32, _vp, _i64, _i32, _vp]),
    ("khg_pdf_posteriors", _i32, [_vp, _i32, _vp, _i64, _i32, _vp, _vp]),
    ("khg_stats_create", _i32, [_vp, _u16, C.POINTER(_vp)]),
    ("khg_stats_zero", _i32, [_vp]),
    ("khg_stats_flags", _i32, [_vp, C.POINTER(_u16)]),
    ("khg_stats_device_buffer", _i32, [_vp, C.POINTER(_vp), C.POINTER(_i64)]),
    ("khg_stats_download", _i32, [_vp, _vp, _vp, _vp, _vp]),
    ("khg_stats_upload", _i32, [_vp, _vp, _vp, _vp, _vp]),
    ("khg_stats_add", _i32, [_vp, _f32, _vp]),
    ("khg_stats_scale", _i32, [_vp, _f32]),
    ("khg_stats_destroy", None, [_vp]),
    ("khg_acc_stats_ali", _i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, C.POINTER(C.c_double)]),
    ("khg_acc_stats_ali_tids", _i32, [_vp, _vp, _vp, _i64, _vp, _vp, _i32, _vp, C.POINTER(C.c_double)]),
    ("khg_acc_from_posteriors", _i32, [_vp, _vp, _i32, _vp, _i64, _i32, _vp]),
    ("khg_estep", _i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _i64, _i64, C.POINTER(C.c_double)]),
    ("khg_mle_update", _i32, [_vp, _vp, _vp, _u16, C.POINTER(_vp), C.POINTER(_f32), C.POINTER(_f32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    ("khg_model_download", _i32, [_vp, _vp, _vp, _vp, _vp, _vp]),
    ("khg_align_batch", _i32, [_vp, _vp, _vp, _i32, _vp, _i32, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    ("khg_align_last_exact_count", _i64, []),
    ("khg_align_last_tile_fraction", C.c_double, []),
    ("khg_align_last_prep_cached", _i32, []),
    ("khg_align_utterance_host", _i32, [_vp, _i32, _vp, _i64, _vp, _i32, _f32, _f32, _f32, _vp, C.POINTER(_i32), C.POINTER(_f32), _vp, _i32, C.POINTER(_i32)]),
    ("khg_gaussian_selection", _i32, [_vp, _i32, _vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _vp, C.POINTER(C.c_double)]),
    ("khg_model_split_by_count", _i32, [_vp, _vp, _i32, _f32, _f32, _f32, _vp, _i64, C.c_uint64, C.POINTER(_vp), C.POINTER(_i32)]),
    ("khg_model_merge_by_count", _i32, [_vp, _vp, _i32, _f32, _f32, C.POINTER(_vp), C.POINTER(_i32)]),
    ("khg_launch_count", _i64, []),
]
KHG_ALIGN_OK, KHG_ALIGN_RETRIED, KHG_ALIGN_FAILED = 0, 1, 2


class GraphBatch(C.Structure):
    """khg_graph_batch of include/khg_b200.h (all HOST arrays)."""
    _fields_ = [("n_utts", _i32), ("frame_offsets", _vp), ("state_offsets", _vp), ("arc_offsets", _vp), ("arc_ilabel", _vp),
                ("arc_nextstate", _vp), ("arc_weight", _vp), ("start_state", _vp), ("final_cost", _vp)]

_lib = None


def lib() -> C.CDLL:
    """Loads libkhg_b200.so; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(kaldi-hmm-gmm_b200/csrc/Makefile). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != KHG_OK:
        msg = lib().khg_last_error()
        raise RuntimeError((msg or b"").decode() or f"khg status {status}")


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def ptr(x, dtype=None):
    """(pointer, loc) of a numpy array / torch tensor / None."""
    if x is None:
        return None, KHG_HOST
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        if dtype is not None and str(x.dtype).split(".")[-1] != np.dtype(dtype).name:
            raise TypeError(f"expected {np.dtype(dtype).name}, got {x.dtype}")
        return x.data_ptr(), (KHG_DEVICE if x.is_cuda else KHG_HOST)
    if not isinstance(x, np.ndarray):
        raise TypeError("expected numpy array or torch tensor")
    if dtype is not None and x.dtype != np.dtype(dtype):
        raise TypeError(f"expected {np.dtype(dtype).name}, got {x.dtype}")
    if not x.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return x.ctypes.data, KHG_HOST


def same_loc(*locs):
    s = {l for l in locs if l is not None}
    if len(s) > 1:
        raise ValueError("all buffers of one call must live on the same side (host or device)")
    return s.pop() if s else KHG_HOST
