"""ctypes binding of libxm3d.so (the C ABI declared in include/xm3d.h).

The library is built in-tree by `make -C xmask3d_b200/csrc` (or `__graft_entry__.build()`).
There is no CPU fallback: a missing library is an ImportError-grade failure, and every
compute call needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("XM3D_SO") or os.path.join(_HERE, "libxm3d.so")   # XM3D_SO: A/B builds of the same ABI
CSRC = os.path.join(_HERE, "csrc")

XM3D_OK = 0
FLAG_VIS_OVERFLOW, FLAG_PAIR_OVERFLOW, FLAG_GRID_RANGE, FLAG_KEY_SENTINEL, FLAG_I16_RANGE, FLAG_VOX_FALLBACK = 1, 2, 4, 8, 16, 32
FLAG_NONFINITE = 64
DEPTH_NONE, DEPTH_U16, DEPTH_F64 = 0, 1, 2
THR_GE_HALF, THR_SIGMOID_GE_HALF, THR_SIGMOID_GT_HALF = 0, 1, 2
MASK_U8, MASK_F32 = 0, 1
VOX_AUTO, VOX_MULTI_KERNEL, VOX_FAST_ONLY = 0, 1, 2
POOL_AUTO, POOL_PAIR_LISTS, POOL_ROWS, POOL_MMA = 0, 1, 2, 3


class View(C.Structure):
    """xm3d_view_t — 192 bytes."""
    _fields_ = [("w2c", C.c_double * 12), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double),
                ("cy", C.c_double), ("pt_off", C.c_int64), ("out_off", C.c_int64), ("depth_off", C.c_int64),
                ("n_pts", C.c_int32), ("depth_h", C.c_int32), ("depth_w", C.c_int32), ("reserved0", C.c_int32),
                ("reserved1", C.c_int64 * 3)]


assert C.sizeof(View) == 192

_P, _I32, _I64, _F64, _F32, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/xm3d.h one to one
PROTOTYPES = {
    "xm3d_version": (C.c_int, []),
    "xm3d_last_error": (C.c_char_p, []),
    "xm3d_device_info": (C.c_int, [_P, _P, _P]),
    "xm3d_launch_count": (_I64, []),
    "xm3d_set_pool_events": (None, [_P, _P]),
    "xm3d_project_ws_bytes": (_SZ, [_I32, _I64, _I32]),
    "xm3d_project_batch": (C.c_int, [_P, _P, _P, _I32, _I64, _P, _I32, _F64, _I32, _I32, _I32, _F64,
                                     _P, _P, _P, _P, _I64, _P, _P, _P, _P, _SZ, _P, _P]),
    "xm3d_unique_ws_bytes": (_SZ, [_I32, _I64]),
    "xm3d_unique_batch": (C.c_int, [_P, _P, _I32, _I64, _P, _P, _P, _P, _P, _I32, _I32, _P, _SZ, _P, _P]),
    "xm3d_fnv_hash_f64": (C.c_int, [_P, _I64, _I32, _P, _P]),
    "xm3d_ravel_hash_f64": (C.c_int, [_P, _I64, _I32, _P, _P, _SZ, _P]),
    "xm3d_ravel_ws_bytes": (_SZ, [_I32]),
    "xm3d_voxelize_ws_bytes": (_SZ, [_I32, _I64]),
    "xm3d_voxel_path_info": (C.c_int, [_P, _I32, _I64, _P, _P]),
    "xm3d_voxelize_batch": (C.c_int, [_P, _I32, _P, _I32, _I64, _P, _P, _P, _P, _P, _I32, _P, _P, _I32, _P, _SZ, _P, _P]),
    "xm3d_mask_words": (_I32, [_I32]),
    "xm3d_gather_ws_bytes": (_SZ, [_I32, _I32, _I32, _I32]),
    "xm3d_gather_masks_batch": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _I64, _P, _P, _P, _SZ, _P]),
    "xm3d_pixel_bits_batch": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "xm3d_pool_ws_bytes": (_SZ, [_I32, _I32, _I32, _I64, _I64]),
    "xm3d_pool_batch": (C.c_int, [_P, _I32, _P, _P, _P, _I32, _I32, _P, _I64, _I64, _I32, _P, _P, _P, _P, _SZ, _P, _P]),
    "xm3d_contra_ws_bytes": (_SZ, [_I32, _I32]),
    "xm3d_contra_select_batch": (C.c_int, [_P, _I32, _P, _P, _I32, _I64, _P, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "xm3d_scatter_batch": (C.c_int, [_P, _P, _I32, _I32, _P, _I64, _P, _I32, _P, _P, _P]),
    "xm3d_point_logits_ws_bytes": (_SZ, [_I32, _I32]),
    "xm3d_point_logits": (C.c_int, [_P, _I64, _I32, _P, _I32, _F32, _P, _P, _P, _P, _I32, _F32, _F32, _P, _P, _P, _SZ, _P]),
    "xm3d_vote_batch": (C.c_int, [_P, _P, _I32, _I64, _P, _P, _I32, _P, _P, _P]),
    "xm3d_vote_argmax": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "xm3d_nn_fill_ws_bytes": (_SZ, [_I32, _I64]),
    "xm3d_nn_fill_batch": (C.c_int, [_P, _P, _P, _I32, _I64, _P, _P, _SZ, _P]),
    "xm3d_segment_max": (C.c_int, [_P, _P, _I32, _I32, _P, _P]),
    "xm3d_mask_prep_batch": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _I32, _P, _P, _P, _P, _P]),
    "xm3d_point_bits_batch": (C.c_int, [_P, _I32, _I32, _I32, _I32, _P, _P, _I64, _P, _P, _P]),
    "xm3d_gather_labels_batch": (C.c_int, [_P, _I32, _I32, _I32, _P, _P, _I64, _P, _P]),
    "xm3d_collate_batch": (C.c_int, [_P, _P, _P, _P, _I32, _I64, _P, _P, _P]),
    "xm3d_pack_i16": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _P]),
    "xm3d_logits_ws_bytes": (_SZ, [_I64, _I32, _I32, _I32]),
    "xm3d_logits": (C.c_int, [_P, _I64, _I32, _P, _I32, _P, _P, _I32, _I32, _F32, _P, _P, _P, _SZ, _P]),
}

_lib = None


class Xm3dError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compile libxm3d.so for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", CSRC, "-s"] + (["-B"] if force else [])
    subprocess.check_call(args)
    return SO_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise Xm3dError(f"{SO_PATH} is missing: run `make -C xmask3d_b200/csrc` "
                            "(or __graft_entry__.build()); there is no CPU fallback")
        l = C.CDLL(SO_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != XM3D_OK:
        raise Xm3dError(f"libxm3d error {rc}: {lib().xm3d_last_error().decode()}")
