"""Drop-in for the reference's `dataset/voxelization_utils.py` (numpy in, numpy out) running on
libxm3d: FNV-1 / ravel keys and the np.unique of `sparse_quantize` are computed on the GPU.

Reference: dataset/voxelization_utils.py:6-18 (fnv_hash_vec), :21-35 (ravel_hash_vec),
:38-102 (sparse_quantize).  Same names, argument meaning, return conventions and assertion
behaviour; there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _dev():
    ops._require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _keys_to_numpy(keys: torch.Tensor) -> np.ndarray:
    return keys.cpu().numpy().view(np.uint64)


def fnv_hash_vec(arr):
    """FNV64-1 of every row, words taken as uint64 (reference :6-18)."""
    assert arr.ndim == 2
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(_dev())
    return _keys_to_numpy(ops.fnv_hash(t))


def ravel_hash_vec(arr):
    """Mixed-radix ravel of the min-shifted rows (reference :21-35)."""
    assert arr.ndim == 2
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(_dev())
    return _keys_to_numpy(ops.ravel_hash(t))


def sparse_quantize(coords, feats=None, labels=None, ignore_label=255,
                    set_ignore_label_when_collision=False, return_index=False,
                    hash_type="fnv", quantization_size=1):
    """Reference :38-102.  Returns exactly what the reference returns for every flag combination."""
    use_label = labels is not None
    use_feat = feats is not None
    if not use_label and not use_feat:
        return_index = True
    assert hash_type in ["ravel", "fnv"], \
        "Invalid hash_type. Either ravel, or fnv allowed. You put hash_type=" + hash_type
    assert coords.ndim == 2, \
        "The coordinates must be a 2D matrix. The shape of the input is " + str(coords.shape)
    if use_feat:
        assert feats.ndim == 2
        assert coords.shape[0] == feats.shape[0]
    if use_label:
        assert coords.shape[0] == len(labels)
    dimension = coords.shape[1]
    if isinstance(quantization_size, (list, tuple, np.ndarray)) or \
            (hasattr(quantization_size, "__len__") and not np.isscalar(quantization_size)):
        assert len(quantization_size) == dimension, "Quantization size and coordinates size mismatch."
        quantization_size = [i for i in quantization_size]
    elif np.isscalar(quantization_size):
        quantization_size = [quantization_size for _ in range(dimension)]
    else:
        raise ValueError("Not supported type for quantization_size.")
    # host side, exactly the reference's arithmetic (float64 divide + floor); the hashing and the
    # unique / inverse maps run on the GPU
    discrete = np.floor(coords / np.array(quantization_size))
    dev = _dev()
    d = torch.from_numpy(np.ascontiguousarray(discrete, dtype=np.float64)).to(dev)
    keys = ops.ravel_hash(d) if hash_type == "ravel" else ops.fnv_hash(d)
    n = discrete.shape[0]
    seg = torch.tensor([0, n], dtype=torch.int64, device=dev)
    u = ops.unique_batch(keys, seg, want_counts=use_label)
    m = int(u.m[0].item())
    inds = u.first[:m].cpu().numpy().astype(np.int64)
    if use_label:
        counts = u.counts[:m].cpu().numpy()
        filtered = labels[inds]
        if set_ignore_label_when_collision:
            filtered[counts > 1] = ignore_label
        if return_index:
            return inds, filtered
        return discrete[inds], feats[inds], filtered
    inverse = u.inverse[:n].cpu().numpy().astype(np.int64)
    if return_index:
        return inds, inverse
    if use_feat:
        return discrete[inds], feats[inds]
    return discrete[inds]
