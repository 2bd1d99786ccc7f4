"""ctypes binding of libdpl_b200.so (the C-ABI declared in include/dpl_b200.h).

The library is the product: there is no Python/torch/CPU fallback behind these
wrappers. If the shared object is missing the import of any kernel wrapper raises
(`DplLibraryError`) with the build command to run.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DPL_LIB: an alternative build of the same library (kernel-tuning experiments), else the in-tree one
LIB_PATH = os.environ.get("DPL_LIB") or os.path.join(_HERE, "libdpl_b200.so")

DPL_SEG_TILE = 8192
DPL_FLAT_TILE = 8192
BLOB_FIELDS = 8  # uint64 fields per dpl_blob

F_PTR, F_NSEG, F_SEGLEN, F_SEG_OUT_BASE, F_SEG_TILE_BEGIN, F_FLAT_TILE_BEGIN, F_STAT, F_RESERVED = range(8)


class DplLibraryError(RuntimeError):
    pass


_lib = None

_c_u64 = ctypes.c_uint64
_c_vp = ctypes.c_void_p
_c_int = ctypes.c_int
_c_size = ctypes.c_size_t
_c_dbl = ctypes.c_double
_c_flt = ctypes.c_float
_c_u32 = ctypes.c_uint32

# name -> (restype, argtypes); must list every symbol of include/dpl_b200.h
SIGNATURES = {
    "dpl_version": (_c_int, []),
    "dpl_last_error": (ctypes.c_char_p, []),
    "dpl_plan_blobs": (_c_int, [_c_vp, _c_int, ctypes.POINTER(_c_u64), ctypes.POINTER(_c_u64),
                                ctypes.POINTER(_c_u64)]),
    "dpl_segstats_scratch_bytes": (_c_size, [_c_u64]),
    "dpl_segstats_f32": (_c_int, [_c_vp, _c_int, _c_u64, _c_u64, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp,
                                  _c_vp, _c_vp, _c_size, _c_int, _c_vp]),
    "dpl_absmax_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_vp]),
    "dpl_hist_abs_f32": (_c_int, [_c_vp, _c_int, _c_u64, _c_vp, _c_int, _c_vp, _c_int, _c_vp]),
    "dpl_hist_percentile": (_c_int, [_c_vp, _c_int, _c_int, _c_dbl, _c_vp, _c_vp, _c_vp, _c_vp,
                                     _c_vp, _c_vp]),
    "dpl_octav_scratch_bytes": (_c_size, [_c_u64, _c_u64]),
    "dpl_octav_f32": (_c_int, [_c_vp, _c_int, _c_u64, _c_u64, _c_vp, _c_vp, _c_dbl, _c_int, _c_vp,
                               _c_vp, _c_vp, _c_size, _c_vp]),
    "dpl_fakequant_f32": (_c_int, [_c_vp, _c_vp, _c_u64, _c_vp, _c_vp, _c_int, _c_u64, _c_int,
                                   _c_int, _c_flt, _c_u64, _c_vp]),
    "dpl_channel_sumdiff_f32": (_c_int, [_c_vp, _c_vp, _c_u64, _c_u64, _c_u64, _c_vp, _c_vp]),
    "dpl_cosine3_f32": (_c_int, [_c_vp, _c_vp, _c_u64, _c_u64, _c_vp, _c_vp]),
    "dpl_adaround_init_f32": (_c_int, [_c_vp, _c_vp, _c_int, _c_u64, _c_vp, _c_vp, _c_vp]),
    "dpl_adaround_weight_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_u64, _c_flt, _c_flt,
                                         _c_int, _c_vp, _c_vp]),
    "dpl_adaround_step_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_u64, _c_flt, _c_flt, _c_flt,
                                       _c_flt, _c_flt, _c_flt, _c_flt, _c_flt, _c_int, _c_flt,
                                       _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "dpl_adaround_step_peer_f32": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_u32, _c_vp, _c_vp, _c_int, _c_u64,
                                            _c_flt, _c_flt, _c_flt, _c_flt, _c_flt, _c_flt, _c_flt, _c_flt,
                                            _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "dpl_recon_act_f32": (_c_int, [_c_vp, _c_vp, _c_u64, _c_int, _c_int, _c_flt, _c_flt, _c_flt, _c_flt,
                                   _c_u64, _c_vp, _c_vp]),
    "dpl_recon_act_bwd_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_u64, _c_int, _c_int, _c_flt, _c_flt,
                                       _c_flt, _c_flt, _c_u64, _c_vp, _c_vp]),
    "dpl_recon_loss_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_u64, _c_int, _c_int, _c_flt, _c_flt, _c_flt,
                                    _c_flt, _c_u64, _c_flt, _c_vp, _c_vp, _c_vp]),
    "dpl_recon_schedule": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_dbl, _c_dbl, _c_dbl, _c_dbl, _c_dbl,
                                    _c_dbl, _c_u64, _c_vp]),
    "dpl_gemm_tf32": (_c_int, [_c_vp, _c_int, ctypes.c_longlong, ctypes.c_longlong, _c_vp, _c_int,
                               ctypes.c_longlong, ctypes.c_longlong, _c_vp, ctypes.c_longlong,
                               ctypes.c_longlong, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp,
                               _c_int, _c_int, _c_vp, _c_vp]),
    "dpl_tf32_residual_f32": (_c_int, [_c_vp, _c_vp, _c_u64, _c_vp]),
    "dpl_tf32_split_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_u64, _c_vp]),
    "dpl_gemm_tf32x3": (_c_int, [_c_vp, _c_vp, _c_int, ctypes.c_longlong, ctypes.c_longlong, _c_vp, _c_int,
                                 ctypes.c_longlong, ctypes.c_longlong, _c_vp, ctypes.c_longlong,
                                 ctypes.c_longlong, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_int, _c_int,
                                 _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "dpl_pad_plane_f32": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                   _c_int, _c_vp]),
    "dpl_conv_taps_tf32x3": (_c_int, [_c_vp, ctypes.c_longlong, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int,
                                      _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_int,
                                      _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "dpl_mix_drop_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_u64, _c_flt, _c_u64, _c_vp]),
    "dpl_conv_direct_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                     _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp,
                                     _c_vp]),
    "dpl_dwconv2d_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                  _c_int, _c_int, _c_vp, _c_vp, _c_vp]),
    "dpl_im2col_f32": (_c_int, [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                _c_int, _c_int, _c_int, _c_vp]),
    "dpl_conv1x1_px_tf32x3": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp,
                                       _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "dpl_clip_f32": (_c_int, [_c_vp, _c_vp, _c_u64, _c_flt, _c_flt, _c_vp, _c_vp, _c_vp]),
    "dpl_add_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_vp, _c_u64, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "dpl_maxpool2d_f32": (_c_int, [_c_vp, _c_vp, _c_u64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                   _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp]),
    "dpl_global_avgpool_f32": (_c_int, [_c_vp, _c_vp, _c_u64, _c_u64, _c_vp, _c_vp, _c_vp]),
    "dpl_tap_conv_tf32": (_c_int, [_c_vp, ctypes.c_longlong, _c_vp, _c_int, _c_vp, _c_int, _c_int, _c_int, _c_int,
                                   _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp,
                                   _c_vp, _c_vp, _c_vp]),
    "dpl_tap_wgrad_tf32": (_c_int, [_c_vp, ctypes.c_longlong, _c_vp, ctypes.c_longlong, _c_vp, _c_int, _c_int,
                                    _c_int, _c_int, _c_vp, _c_vp, _c_vp, _c_vp]),
    "dpl_taps_layout_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp]),
    "dpl_dwconv2d_wgrad_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                        _c_int, _c_int, _c_vp]),
    "dpl_dwconv2d_dgrad_f32": (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                        _c_int, _c_int, _c_vp]),
}


def lib():
    """Load (once) and return the ctypes handle; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DplLibraryError(
            f"{LIB_PATH} is not built. Run `python -m dipoorlet_b200.build` (needs nvcc; "
            "cross-compiles for sm_100a without a GPU). There is no CPU fallback.")
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as e:  # stale build
            raise DplLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    if handle.dpl_version() != 100:
        raise DplLibraryError("libdpl_b200.so version mismatch; rebuild it")
    _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().dpl_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with status {status}: {msg}")


def plan_blobs(table):
    """Fill the *_begin fields of a host blob table (uint64 [n, 8]) in place.

    Returns (n_segments, n_seg_tiles, n_flat_tiles)."""
    assert table.dtype == np.uint64 and table.ndim == 2 and table.shape[1] == BLOB_FIELDS
    assert table.flags["C_CONTIGUOUS"]
    a, b, c = _c_u64(), _c_u64(), _c_u64()
    check(lib().dpl_plan_blobs(table.ctypes.data, table.shape[0], ctypes.byref(a), ctypes.byref(b),
                               ctypes.byref(c)), "dpl_plan_blobs")
    return a.value, b.value, c.value


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
