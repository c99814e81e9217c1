"""ORACLE — ctypes front-end of oracle/xm3d_oracle.c (test infrastructure, NOT product code)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libxm3d_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "xm3d_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.xo_project.restype = C.c_int
        _lib.xo_unique_u64.restype = C.c_int
        _lib.xo_voxelize.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def project(xyz, w2c, intr, depth=None, depth_scale=1000.0, image_dim=(320, 240), cut=10, vis_thres=0.25):
    """depth: None | uint16 [h,w] (raw, divided by depth_scale) | float64 [h,w] metres."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    w2c = np.ascontiguousarray(w2c, np.float64)
    n = xyz.shape[0]
    out = np.zeros((n, 3), np.int64)
    kind, dh, dw, dp = 0, 0, 0, None
    if depth is not None:
        if depth.dtype == np.uint16:
            kind, dp = 1, np.ascontiguousarray(depth)
        else:
            kind, dp = 2, np.ascontiguousarray(depth, np.float64)
        dh, dw = dp.shape
    fx, fy, cx, cy = intr
    rc = lib().xo_project(_p(xyz), C.c_int64(n), _p(w2c), C.c_double(fx), C.c_double(fy), C.c_double(cx),
                          C.c_double(cy), _p(dp), kind, dh, dw, C.c_double(depth_scale),
                          int(image_dim[0]), int(image_dim[1]), int(cut), C.c_double(vis_thres), _p(out))
    assert rc == 0
    return out


def fnv(coords):
    coords = np.ascontiguousarray(coords, np.float64)
    out = np.zeros(coords.shape[0], np.uint64)
    lib().xo_fnv(_p(coords), C.c_int64(coords.shape[0]), int(coords.shape[1]), _p(out))
    return out


def unique_u64(keys):
    keys = np.ascontiguousarray(keys, np.uint64)
    n = keys.shape[0]
    first, inv, cnt = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64)
    m = C.c_int64(0)
    rc = lib().xo_unique_u64(_p(keys), C.c_int64(n), _p(first), _p(inv), _p(cnt), C.byref(m))
    assert rc == 0
    return first[:m.value].copy(), inv, cnt[:m.value].copy()


def voxelize(xyz, rt):
    xyz = np.ascontiguousarray(xyz, np.float32)
    rt = np.ascontiguousarray(rt, np.float64)
    n = xyz.shape[0]
    first, inv = np.zeros(n, np.int64), np.zeros(n, np.int64)
    vox = np.zeros((n, 3), np.float64)
    m = C.c_int64(0)
    rc = lib().xo_voxelize(_p(xyz), C.c_int64(n), _p(rt), _p(first), _p(inv), _p(vox), C.byref(m))
    assert rc == 0
    return vox[:m.value].copy(), first[:m.value].copy(), inv


def pool_label_f64(feat, label, k):
    feat = np.ascontiguousarray(feat, np.float32)
    label = np.ascontiguousarray(label, np.int32)
    n, c = feat.shape
    s, cnt = np.zeros((k, c), np.float64), np.zeros(k, np.int64)
    lib().xo_pool_label_f64(_p(feat), C.c_int64(n), int(c), _p(label), int(k), _p(s), _p(cnt))
    return s, cnt


def pool_member_f64(feat, member):
    feat = np.ascontiguousarray(feat, np.float32)
    member = np.ascontiguousarray(member, np.uint8)
    (n, c), k = feat.shape, member.shape[0]
    s, cnt = np.zeros((k, c), np.float64), np.zeros(k, np.int64)
    lib().xo_pool_member_f64(_p(feat), C.c_int64(n), int(c), _p(member), int(k), _p(s), _p(cnt))
    return s, cnt


def scatter_member_f32(member, emb):
    member = np.ascontiguousarray(member, np.uint8)
    emb = np.ascontiguousarray(emb, np.float32)
    (k, n), c = member.shape, emb.shape[1]
    out, counter = np.zeros((n, c), np.float32), np.zeros(n, np.float32)
    lib().xo_scatter_member_f32(_p(member), int(k), C.c_int64(n), _p(emb), int(c), _p(out), _p(counter))
    return out, counter
