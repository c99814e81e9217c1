"""ORACLE — test infrastructure, NOT product code.

The reference arm of the benchmark: ONE (scene, view) unit of the hot path on the host CPU, run with the
reference's own functions when `oracle/_ref/` holds them (vendored unmodified by oracle/vendor_ref.py; kind
"reference"), else with the numpy / torch-CPU port (oracle/ref_port.py; kind "port").  The pooling loop has no
named function in the reference (it is inline at models/utils/criterion.py:83-94, 148-157), so it is the port's
restatement in both cases.  Only bench.py's CPU legs and tests may import this module.
"""
from __future__ import annotations

import collections
import collections.abc
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_impl = None


def available() -> bool:
    return os.path.exists(os.path.join(_REF, "MANIFEST.json"))


def load():
    """(kind, Voxelizer class, getMapping function).  Cached."""
    global _impl
    if _impl is None:
        from . import ref_port as P
        if available():
            collections.Sequence = collections.abc.Sequence        # the reference targets Python 3.9
            collections.Iterable = collections.abc.Iterable
            if _REF not in sys.path:
                sys.path.insert(0, _REF)
            from dataset.voxelizer import Voxelizer                # noqa: E402  (oracle/_ref/dataset/voxelizer.py)
            from models.utils.mapping_util import getMapping       # noqa: E402
            _impl = ("reference", Voxelizer, getMapping)
        else:
            _impl = ("port", P.Voxelizer, P.getMapping)
    return _impl


def one_view(mapper, voxelizer_cls, vox_kwargs, seed, scene_xyz, scene_colors, scene_labels, pose, depth_m, mask_f32,
             feat_buf):
    """The loader + pooling work of one (scene, view): compute_mapping -> compaction (dataset/data_loader_infer.py:
    174-182) -> Voxelizer.voxelize (np.random seeded as the harness prescribes) -> mask-at-point gather + per-mask
    `feature[mask].mean(0)` (criterion.py:83-85, 152-157).  Returns the number of visible points."""
    import torch
    from . import ref_port as P
    mapping = mapper.compute_mapping(pose, scene_xyz, depth_m)
    vis, x_label, y_label = P.compact_mapping(mapping)
    locs, col, lab = scene_xyz[vis], scene_colors[vis], scene_labels[vis]
    n = locs.shape[0]
    if n == 0:
        return 0
    np.random.seed(seed)
    voxelizer_cls(**vox_kwargs).voxelize(locs, col, lab)
    member = P.gather_masks(mask_f32, torch.from_numpy(x_label), torch.from_numpy(y_label), "ge0.5")
    P.masked_mean_pool(feat_buf[:n], member)
    return n
