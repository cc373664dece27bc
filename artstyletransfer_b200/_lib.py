"""ctypes binding of libast_sm100.so (the stub a reference maintainer would add — see INTEGRATION.md).

Every function is declared exactly as in include/ast_sm100.h.  `call()` turns a negative status into a
RuntimeError carrying ast_last_error().  The library is looked up next to this file only (in-tree build:
`python artstyletransfer_b200/build.py`); a missing library is an ImportError, never a fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libast_sm100.so')

AST_ABI_VERSION = 9
AST_PREC_TF32, AST_PREC_FP32, AST_PREC_BF16 = 0, 1, 2
AST_LAYOUT_CHW, AST_LAYOUT_HWC = 0, 1
AST_COORD_TORCH, AST_COORD_CV2 = 0, 1
AST_INIT_RANDOM, AST_INIT_CONTENT_NOISE = 0, 1
AST_NOISE_MAX_LEVELS = 16

PRECISIONS = {'tf32': AST_PREC_TF32, 'fp32': AST_PREC_FP32, 'bf16': AST_PREC_BF16}


class NoiseLevel(C.Structure):
    """struct ast_noise_level (include/ast_sm100.h)."""
    _fields_ = [('lowres', C.c_void_p), ('lh', C.c_int32), ('lw', C.c_int32), ('kind', C.c_int32),
                ('pad_', C.c_int32), ('gy', C.c_void_p), ('gx', C.c_void_p), ('center', C.c_double),
                ('central', C.c_double), ('peripheral', C.c_double)]


class HaloRow(C.Structure):
    """struct ast_halo_row (include/ast_sm100.h)."""
    _fields_ = [('src', C.c_void_p), ('dst_remote', C.c_void_p), ('stage', C.c_void_p), ('halo', C.c_void_p),
                ('flag_remote', C.c_void_p), ('flag_local', C.c_void_p), ('state', C.c_void_p),
                ('bytes', C.c_int64), ('slot_stride', C.c_int64)]


class FinalizeItem(C.Structure):
    """struct ast_finalize_item (include/ast_sm100.h)."""
    _fields_ = [('G_raw', C.c_void_p), ('A', C.c_void_p), ('out', C.c_void_p), ('loss', C.c_void_p),
                ('scale', C.c_float), ('C', C.c_int32), ('round_out', C.c_int32), ('pad_', C.c_int32)]


AST_HALO_MAX_ROWS = 16
AST_FINALIZE_MAX_ITEMS = 32
AST_GATHER_MAX_PEERS = 7
AST_GATHER_MAX_SEGS = 24


class BandGatherDesc(C.Structure):
    """struct ast_band_gather_desc (include/ast_sm100.h)."""
    _fields_ = [('local_base', C.c_void_p), ('peer_base', C.c_void_p * AST_GATHER_MAX_PEERS),
                ('ready_remote', C.c_void_p * AST_GATHER_MAX_PEERS), ('ready_local', C.c_void_p * AST_GATHER_MAX_PEERS),
                ('arrive_remote', C.c_void_p * AST_GATHER_MAX_PEERS), ('arrive_local', C.c_void_p * AST_GATHER_MAX_PEERS),
                ('state', C.c_void_p), ('seg_off', C.c_int64 * AST_GATHER_MAX_SEGS),
                ('seg_bytes', C.c_int64 * AST_GATHER_MAX_SEGS), ('n_peers', C.c_int32), ('n_segs', C.c_int32)]
_p, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); one entry per declaration in include/ast_sm100.h
SIGNATURES = {
    'ast_version': (_i, []),
    'ast_last_error': (C.c_char_p, []),
    'ast_device_check': (_i, []),
    'ast_gram_workspace_bytes': (_sz, [_i, _i64]),
    'ast_gram_tf32_supported': (_i, [_p, _i, _i64, _i64]),
    'ast_gram_mse_fwd': (_i, [_p, _i, _i64, _i64, _f, _p, _p, _p, _p, _sz, _i, _p]),
    'ast_gram_finalize': (_i, [_p, _i, _f, _p, _p, _p, _p, _sz, _i, _p]),
    'ast_gram_bwd': (_i, [_p, _p, _i, _i64, _i64, _f, _p, _p, _i, _i, _p]),
    'ast_finalize_batch_workspace_bytes': (_sz, [_i]),
    'ast_gram_finalize_batch': (_i, [C.POINTER(FinalizeItem), _i, _p, _sz, _p]),
    'ast_gram_mse_fwd_nhwc': (_i, [_p, _i, _i64, _f, _p, _p, _p, _p, _sz, _i, _p]),
    'ast_gram_bwd_nhwc': (_i, [_p, _p, _i, _i64, _f, _p, _p, _i, _i, _i, _p]),
    'ast_gram_bwd_nhwc_bf16': (_i, [_p, _p, _i, _i64, _f, _p, _p, _i, _i, _p]),
    'ast_reduce_workspace_bytes': (_sz, []),
    'ast_mse_fwd': (_i, [_p, _p, _i64, _f, _p, _p, _sz, _p]),
    'ast_mse_bwd': (_i, [_p, _p, _i64, _f, _p, _p, _i, _i, _p]),
    'ast_bias_relu_nhwc': (_i, [_p, _p, _i, _i64, _p]),
    'ast_relu_bwd': (_i, [_p, _p, _i64, _p]),
    'ast_maxpool2x2_nhwc': (_i, [_p, _i, _i, _i, _p, _p]),
    'ast_maxpool2x2_bwd_nhwc': (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    'ast_chw_to_hwc': (_i, [_p, _i, _i64, _i64, _p, _p]),
    'ast_hwc_to_chw': (_i, [_p, _i, _i64, _p, _i64, _i, _p]),
    'ast_unprepare_hwc': (_i, [_p, _i64, _d, _d, _d, _p, _p]),
    'ast_tv_fwd': (_i, [_p, _i, _i, _i, _p, _p, _p, _sz, _p]),
    'ast_tv_bwd': (_i, [_p, _i, _i, _i, _p, _f, _f, _p, _p, _i, _p]),
    'ast_tv_bwd_rows': (_i, [_p, _i, _i, _i, _i, _i, _p, _f, _f, _p, _p, _i, _p]),
    'ast_level_combine': (_i, [_p, _i, _p, _p, _f, _f, _f, _p, _p]),
    'ast_bicubic_down2x': (_i, [_p, _i, _i, _i, _p, _p]),
    'ast_bicubic_down2x_adj': (_i, [_p, _i, _i, _i, _p, _i, _p]),
    'ast_bicubic_down2x_tv_workspace_bytes': (_sz, [_i, _i, _i]),
    'ast_bicubic_down2x_tv': (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    'ast_bicubic_resize': (_i, [_p, _i, _i, _i, _p, _i, _i, _i, _i, _p]),
    'ast_bicubic_resize_adj': (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _i, _p]),
    'ast_noise_init': (_i, [_p, _i, _i, C.POINTER(NoiseLevel), _i, _d, _i, _i, _d, _d, _p, _p]),
    'ast_halo_exchange': (_i, [C.POINTER(HaloRow), _i, _p]),
    'ast_band_announce': (_i, [C.POINTER(BandGatherDesc), _p]),
    'ast_band_gather': (_i, [C.POINTER(BandGatherDesc), _p]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} not found: build it with `python artstyletransfer_b200/build.py` '
                '(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.')
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.ast_version() != AST_ABI_VERSION:
            raise ImportError(f'libast_sm100.so ABI {lib.ast_version()} != expected {AST_ABI_VERSION}')
        _lib = lib
    return _lib


def last_error() -> str:
    return load().ast_last_error().decode('utf-8', 'replace')


def call(name: str, *args):
    """Invoke an int-returning entry point; raise RuntimeError(ast_last_error()) on failure."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError(f'{name} failed ({rc}): {last_error()}')
    return rc
