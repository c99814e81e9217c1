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


def _validate(coords, feats, labels, hash_type):
    """Same failure behaviour as the reference (:53-66): AssertionError on a bad hash type or bad shapes."""
    if hash_type not in ("ravel", "fnv"):
        raise AssertionError(f"hash_type must be 'ravel' or 'fnv', got {hash_type!r}")
    if coords.ndim != 2:
        raise AssertionError(f"coords must be a 2-D [n, dim] array, got shape {coords.shape}")
    if feats is not None and (feats.ndim != 2 or feats.shape[0] != coords.shape[0]):
        raise AssertionError("feats must be a 2-D array with one row per coordinate")
    if labels is not None and len(labels) != coords.shape[0]:
        raise AssertionError("labels must hold one entry per coordinate")


def _cell_size(quantization_size, dim):
    """Per-axis cell size (:68-77): a scalar is broadcast, a sequence must have `dim` entries."""
    if np.isscalar(quantization_size):
        return np.full(dim, quantization_size)
    if isinstance(quantization_size, (list, tuple, np.ndarray, torch.Tensor)):
        if len(quantization_size) != dim:
            raise AssertionError("quantization_size needs one entry per coordinate axis")
        return np.array([q for q in quantization_size])
    raise ValueError("Not supported type for quantization_size.")


def sparse_quantize(coords, feats=None, labels=None, ignore_label=255,
                    set_ignore_label_when_collision=False, return_index=False,
                    hash_type="fnv", quantization_size=1):
    """Reference :38-102.  Returns exactly what the reference returns for every flag combination:
    with labels (inds, labels') or (cells, feats, labels'); without labels (inds, inverse), (cells, feats)
    or cells — where inds / inverse are np.unique's first-occurrence and inverse maps of the cell keys."""
    _validate(coords, feats, labels, hash_type)
    with_labels, with_feats = labels is not None, feats is not None
    want_index = return_index or not (with_labels or with_feats)
    # host side, the reference's arithmetic (float64 divide + floor, :78); hashing and the unique /
    # first-index / inverse / count maps run on the GPU
    cells = np.floor(coords / _cell_size(quantization_size, coords.shape[1]))
    dev = _dev()
    d = torch.from_numpy(np.ascontiguousarray(cells, dtype=np.float64)).to(dev)
    keys = ops.ravel_hash(d) if hash_type == "ravel" else ops.fnv_hash(d)
    n = cells.shape[0]
    u = ops.unique_batch(keys, torch.tensor([0, n], dtype=torch.int64, device=dev), want_counts=with_labels)
    m = int(u.m[0].item())
    inds = u.first[:m].cpu().numpy().astype(np.int64)
    if with_labels:                                             # :86-93
        kept = labels[inds]
        if set_ignore_label_when_collision:
            kept[u.counts[:m].cpu().numpy() > 1] = ignore_label
        return (inds, kept) if want_index else (cells[inds], feats[inds], kept)
    if want_index:                                              # :95-97
        return inds, u.inverse[:n].cpu().numpy().astype(np.int64)
    return (cells[inds], feats[inds]) if with_feats else cells[inds]
