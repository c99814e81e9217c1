"""CPU-side checks of the drop-in boundary: libxm3d.so loads without a GPU, exports every
symbol include/xm3d.h declares, and the ctypes prototypes cover exactly that set."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from xmask3d_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    return _lib


def _declared():
    src = open(os.path.join(ROOT, "include", "xm3d.h")).read()
    return sorted(set(re.findall(r"XM3D_API[^;(]*?\b(xm3d_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 20
    dll = ctypes.CDLL(lib.SO_PATH)
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/xm3d.h but not exported"
    assert sorted(lib.PROTOTYPES) == names


def test_abi_constants_and_struct(lib):
    l = lib.lib()
    assert l.xm3d_version() == 100
    assert ctypes.sizeof(lib.View) == 192 and lib.View.n_pts.offset == 152 and lib.View.fx.offset == 96
    hdr = open(os.path.join(ROOT, "include", "xm3d.h")).read()
    for name, val in (("XM3D_DEPTH_U16", lib.DEPTH_U16), ("XM3D_DEPTH_F64", lib.DEPTH_F64),
                      ("XM3D_THR_SIGMOID_GE_HALF", lib.THR_SIGMOID_GE_HALF),
                      ("XM3D_THR_SIGMOID_GT_HALF", lib.THR_SIGMOID_GT_HALF), ("XM3D_MASK_F32", lib.MASK_F32),
                      ("XM3D_FLAG_GRID_RANGE", lib.FLAG_GRID_RANGE), ("XM3D_FLAG_KEY_SENTINEL", lib.FLAG_KEY_SENTINEL)):
        assert re.search(rf"#define {name} {val}\b", hdr), name


def test_workspace_queries_need_no_gpu(lib):
    l = lib.lib()
    assert l.xm3d_mask_words(50) == 2 and l.xm3d_mask_words(100) == 4 and l.xm3d_mask_words(32) == 1
    assert l.xm3d_project_ws_bytes(20, 3_000_000, 150_000) > 3_000_000 * 8
    assert l.xm3d_voxelize_ws_bytes(160, 6_400_000) > 6_400_000 * (24 + 20)
    assert l.xm3d_gather_ws_bytes(2, 50, 240, 320) >= 2 * 240 * 320 * 2 * 4
    assert l.xm3d_pool_ws_bytes(160, 50, 768, 6_400_000, 6_400_000) > 6_400_000 * 4 + 160 * 50 * 768 * 4
    # (round 2: only the text side is staged — the mask embeddings are split in shared memory)
    assert 2 * 20 * 768 * 4 < l.xm3d_logits_ws_bytes(8000, 19, 768, 19) < 2 * 8020 * 768 * 4


def test_compute_calls_fail_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    from xmask3d_b200.voxelization_utils import fnv_hash_vec
    with pytest.raises(lib.Xm3dError):
        fnv_hash_vec(np.zeros((4, 3)))
    assert l_device_info_fails(lib)


def l_device_info_fails(lib):
    sm = ctypes.c_int32(0)
    return lib.lib().xm3d_device_info(ctypes.byref(sm), None, None) != 0
